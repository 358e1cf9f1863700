"""CPU: the UNMODIFIED reference with its real main() (main.c refuses one task) on 2 and 4 tasks of this box, through the
mini-MPI of oracle/stubs/minimpi (fork + shared-memory mailboxes).  Pins the mini-MPI on the reference's own parallel code:
read_ic.c distribution, domain.c ORB + particle exchange, the hypercube exchanges of gravtree.c / sidm.c, timeline.c
reductions and the parallel snapshot writer of io.c all have to work for the two runs to end at the same state.  The two
decompositions build different per-task trees (forces differ at the tree's own error), hence tolerances."""
import os

import numpy as np
import pytest

import mpi_case

EXE = os.path.join(mpi_case.ROOT, "oracle", "_ref", "sidm_ref_mpi")


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/sidm_ref_mpi not built (needs /root/reference)")
def test_reference_main_on_two_and_four_tasks(tmp_path):
    runs = {}
    for nt in (2, 4):
        w = str(tmp_path / f"np{nt}")
        mpi_case.write_case(w, 4000, TimeMax=0.005)
        r = mpi_case.run_case(w, "sidm_ref_mpi", nt, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        assert "Number of processors MUST" not in r.stdout
        runs[nt] = mpi_case.last_snapshot(w)
    a, b = runs[2], runs[4]
    assert a["time"] == b["time"] == 0.005
    assert np.array_equal(a["ids"], b["ids"]) and len(a["ids"]) == 4000
    np.testing.assert_allclose(a["pos"], b["pos"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(a["vel"], b["vel"], rtol=0, atol=0.5)
    assert np.abs(a["vel"] - b["vel"]).mean() < 5e-3


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/sidm_ref_mpi not built (needs /root/reference)")
def test_reference_stop_file_and_restart(tmp_path):
    """the reference's own interruption / restart cycle (run.c:152-176, restart.c) on two tasks under the mini-MPI, initial
    conditions and snapshots split over two files: with the tree rebuilt every step (TreeUpdateFrequency 0) the continued run
    ends bit-identical to the uninterrupted one.  The GPU drop-in is held to the same in tests/test_gpu_mpi_dropin.py."""
    import subprocess
    import oracle
    kw = dict(NumFilesPerSnapshot=2, TimeMax=0.006, TreeUpdateFrequency=0.0)

    def final(w):
        files = sorted(f for f in os.listdir(w) if f.startswith("snp_") and f.count(".") == 1)
        last = files[-1].split(".")[0]
        parts = [oracle.read_snapshot(os.path.join(w, f"{last}.{k}")) for k in range(2)]
        ids = np.concatenate([p["ids"] for p in parts]); o = np.argsort(ids)
        return parts[0]["time"], ids[o], np.concatenate([p["pos"] for p in parts])[o], np.concatenate([p["vel"] for p in parts])[o]
    w1 = str(tmp_path / "straight")
    mpi_case.write_case(w1, 4000, **kw)
    assert mpi_case.run_case(w1, "sidm_ref_mpi", 2).returncode == 0
    w2 = str(tmp_path / "restarted")
    mpi_case.write_case(w2, 4000, **kw)
    open(os.path.join(w2, "stop"), "w").close()
    assert mpi_case.run_case(w2, "sidm_ref_mpi", 2).returncode == 0
    assert os.path.exists(os.path.join(w2, "rst_out.1")) and not any(f.startswith("snp_") for f in os.listdir(w2))
    os.remove(os.path.join(w2, "stop"))
    r = subprocess.run([EXE, "param.txt", "1"], cwd=w2, capture_output=True, text=True, timeout=600, env=dict(os.environ, MINIMPI_NP="2"))
    assert r.returncode == 0, r.stdout[-1500:]
    a, b = final(w1), final(w2)
    assert a[0] == b[0] == 0.006 and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
