"""GPU: the multi-task drop-in.  The reference's REAL main() (main.c:39-53 refuses NTask <= 1) with its unmodified driver
(begrun.c, init.c, read_ic.c, run.c, timeline.c, timestep.c, predict.c, domain.c ...) linked against the shim
(-DB200_SHIM_ACCEL) + b200_comm.c + libsidm_b200.so, started on 2 and on 4 tasks through the mini-MPI (one GPU per task
with NCCL when the box has them, else the tasks share GPUs and the all-gather is staged through the host).

* 2 tasks and 4 tasks end in the same state BIT FOR BIT (by particle ID): the result does not depend on how domain.c
  distributes the particles over the tasks, nor on how the library deals the work out over the GPUs.
* against the all-CPU reference on 2 tasks (its per-task trees differ from the global tree at the tree's own error):
  positions to 2e-3 kpc, velocities to 0.5 km/s after the same 0.01 time units."""
import os

import numpy as np
import pytest

import mpi_case

pytestmark = pytest.mark.gpu
N = 20000


def _have(exe):
    return os.path.exists(os.path.join(mpi_case.ROOT, "oracle", "_ref", exe))


@pytest.mark.skipif(not _have("sidm_b200_mpi"), reason="oracle/_ref/sidm_b200_mpi not built (needs /root/reference)")
def test_real_main_on_two_and_four_tasks(tmp_path):
    runs, logs = {}, {}
    for nt in (2, 4):
        w = str(tmp_path / f"gpu{nt}")
        mpi_case.write_case(w, N, TimeMax=0.01)
        # 2 tasks: even this small problem is dealt out over the GPUs (exchanges in every phase); 4 tasks: default threshold
        r = mpi_case.run_case(w, "sidm_b200_mpi", nt, timeout=900, env=dict(B200_SHARD_MIN_WORK="1") if nt == 2 else None)
        assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-2500:]
        assert f"libsidm_b200 on {nt} tasks" in r.stdout
        runs[nt], logs[nt] = mpi_case.last_snapshot(w), r.stdout
    a, b = runs[2], runs[4]
    assert a["time"] == b["time"] == 0.01 and np.array_equal(a["ids"], b["ids"]) and len(a["ids"]) == N
    assert np.array_equal(a["pos"], b["pos"]) and np.array_equal(a["vel"], b["vel"]), "2-task and 4-task runs differ"
    if _have("sidm_ref_mpi"):
        w = str(tmp_path / "cpu2")
        mpi_case.write_case(w, N, TimeMax=0.01)
        r = mpi_case.run_case(w, "sidm_ref_mpi", 2, timeout=900)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        c = mpi_case.last_snapshot(w)
        assert c["time"] == a["time"] and np.array_equal(c["ids"], a["ids"])
        np.testing.assert_allclose(a["pos"], c["pos"], rtol=0, atol=2e-3)
        np.testing.assert_allclose(a["vel"], c["vel"], rtol=0, atol=0.5)
        assert np.abs(a["vel"] - c["vel"]).mean() < 5e-3


@pytest.mark.skipif(not _have("sidm_b200_mpi"), reason="oracle/_ref/sidm_b200_mpi not built (needs /root/reference)")
def test_restart_and_split_snapshot(tmp_path):
    """restart.c through the drop-in, with scatterings (sigma/m = 38 cm^2/g): a run that is interrupted by the reference's
    stop-file mechanism (run.c:152-176: restart files written after the current step) and continued with RestartFlag = 1 ends
    bit-identical to the uninterrupted run - the restart files come from the host arrays the shim keeps current, the
    generator state of the path travels in the .b200rng side file.  Initial conditions and snapshots are split over two
    files (NumFilesPerSnapshot = 2, one per task: read_ic.c:62-75, io.c:78-103)."""
    import subprocess
    import oracle
    kw = dict(CrossSection=38.2614, NumFilesPerSnapshot=2, TimeMax=0.006)
    w1 = str(tmp_path / "straight")
    mpi_case.write_case(w1, N, **kw)
    r = mpi_case.run_case(w1, "sidm_b200_mpi", 2)
    assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-2500:]
    w2 = str(tmp_path / "restarted")
    mpi_case.write_case(w2, N, **kw)
    open(os.path.join(w2, "stop"), "w").close()              # stop after the first step
    r = mpi_case.run_case(w2, "sidm_b200_mpi", 2)
    assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-2500:]
    assert os.path.exists(os.path.join(w2, "rst_out.0")) and os.path.exists(os.path.join(w2, "rst_out.0.b200rng"))
    assert not any(f.startswith("snp_") for f in os.listdir(w2)), "the interrupted run must not have reached TimeMax"
    os.remove(os.path.join(w2, "stop"))
    e = dict(os.environ, MINIMPI_NP="2")
    r = subprocess.run([os.path.join(mpi_case.ROOT, "oracle", "_ref", "sidm_b200_mpi"), "param.txt", "1"], cwd=w2, capture_output=True, text=True, timeout=900, env=e)
    assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-2500:]

    def final(w):
        files = sorted(f for f in os.listdir(w) if f.startswith("snp_") and f.count(".") == 1)
        last = files[-1].split(".")[0]
        parts = [oracle.read_snapshot(os.path.join(w, f"{last}.{k}")) for k in range(2)]
        assert all(p["time"] == 0.006 for p in parts)
        assert all(p["npart"][1] == len(p["ids"]) for p in parts)
        ids = np.concatenate([p["ids"] for p in parts]); o = np.argsort(ids)
        return ids[o], np.concatenate([p["pos"] for p in parts])[o], np.concatenate([p["vel"] for p in parts])[o]
    a, b = final(w1), final(w2)
    assert len(a[0]) == N and np.array_equal(a[0], b[0])
    assert "SCT" in r.stdout and any(int(l.split()[3]) > 0 for l in r.stdout.splitlines() if l.startswith("SCT ")), "no scatterings after the restart"
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), "restarted run differs from the uninterrupted one"
