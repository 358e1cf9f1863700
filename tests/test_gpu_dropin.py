"""GPU: the drop-in boundary.  The reference's own driver code (accel.c, init.c, timeline.c,
predict.c ... compiled unmodified) is linked once against its CPU hot path and once against
sidm-nbody_b200/shim/b200_shim.c + libsidm_b200.so; both run the same start-up and one
compute_accelerations(0); the particle fields the path owns must agree."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
N = 30000


def _run(kind, out):
    r = subprocess.run([sys.executable, os.path.join(HERE, "dropin_runner.py"), kind, out, str(N)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return np.load(out)


@pytest.mark.parametrize("kind", ["b200", "b200f"])
def test_reference_driver_on_gpu_path(tmp_path, refdrv_mod, kind):
    """b200: the reference's accel.c calls the shim's gravity_tree() / sidm() / sidm_ensure_neighbours();
    b200f: accel.c's compute_accelerations() itself is the shim's (-DB200_SHIM_ACCEL, one coarse call)"""
    if not refdrv_mod.available(kind):
        pytest.skip(f"oracle/_ref/libsidmref_{kind}.so not built")
    cpu = _run("diag", str(tmp_path / "cpu.npz"))
    gpu = _run(kind, str(tmp_path / "gpu.npz"))
    rms = lambda a, b: float(np.sqrt(((a.astype(np.float64) - b) ** 2).sum() / (b.astype(np.float64) ** 2).sum()))
    # start-up: smoothing lengths from the reference's own init.c loop over the GPU k-NN / counts
    assert np.array_equal(cpu["ngb0"], gpu["ngb0"])
    assert (cpu["h0"] == gpu["h0"]).mean() > 0.999
    np.testing.assert_allclose(gpu["h0"], cpu["h0"], rtol=3e-7)
    # compute_accelerations(1): BH forces
    assert rms(gpu["acc1"], cpu["acc1"]) < 2e-6
    np.testing.assert_allclose(gpu["old1"], cpu["old1"], rtol=2e-5)
    # compute_accelerations(0): relative criterion with that OldAcc, prediction to t=0.01, repair loop
    assert np.array_equal(cpu["pospred"], gpu["pospred"])
    assert rms(gpu["acc2"], cpu["acc2"]) < 1e-4
    assert np.array_equal(cpu["ngb2"], gpu["ngb2"])
    assert (cpu["h2"] == gpu["h2"]).mean() > 0.999
    assert not gpu["dvel"].any() and not cpu["dvel"].any()
    assert cpu["nactive"] == gpu["nactive"] == N
    if kind == "b200f":                                # compute_potential() through the shim
        np.testing.assert_allclose(gpu["pot"], cpu["pot"], rtol=3e-6)
        # compute_global_quantities_of_system() and savepositions() through the shim (global.c / io.c replaced)
        s_cpu, s_gpu = cpu["sys"], gpu["sys"]
        # (VelPred carries the 2e-6 different start-up accelerations of the two runs: E_kin to 1e-8, not to rounding)
        assert abs(s_gpu[0] - s_cpu[0]) <= 1e-12 * s_cpu[0] and abs(s_gpu[1] - s_cpu[1]) <= 1e-8 * s_cpu[1]       # Mass, EnergyKin
        assert abs(s_gpu[2] - s_cpu[2]) <= 3e-6 * abs(s_cpu[2])                                                      # EnergyPot
        np.testing.assert_allclose(s_gpu, s_cpu, rtol=1e-5, atol=1e-6 * np.abs(s_cpu).max())
        # same file layout; VelPred carries the (2e-6 different) start-up accelerations, everything else is bit-equal
        import oracle
        snaps = {}
        for k, r in (("cpu", cpu), ("gpu", gpu)):
            (tmp_path / f"snap_{k}").write_bytes(r["snap"].tobytes())
            snaps[k] = oracle.read_snapshot(str(tmp_path / f"snap_{k}"))
        assert len(gpu["snap"]) == len(cpu["snap"]) and gpu["snap"][:264].tobytes() == cpu["snap"][:264].tobytes()
        assert np.array_equal(snaps["gpu"]["pos"], snaps["cpu"]["pos"]) and np.array_equal(snaps["gpu"]["ids"], snaps["cpu"]["ids"])
        assert snaps["gpu"]["mass"] is None and snaps["cpu"]["mass"] is None
        np.testing.assert_allclose(snaps["gpu"]["vel"], snaps["cpu"]["vel"], rtol=1e-5, atol=1e-4)


def _run_loop(kind, out):
    r = subprocess.run([sys.executable, os.path.join(HERE, "dropin_runner.py"), kind, out, str(N), "run"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return np.load(out)


@pytest.mark.parametrize("kind", ["b200", "b200f"])
def test_reference_main_loop_individual_timesteps(tmp_path, refdrv_mod, kind):
    """BASELINE config C5 ingredients through the reference's own driver: 60 iterations of the main loop of run.c
    (find_next_time -> compute_accelerations(0) -> advance -> find_timesteps(0), timeline.c / predict.c / timestep.c
    unmodified) with individual time steps, i.e. active sets of 10^2..10^4 of 2e4 particles taken from the ForceFlag
    chain, all-CPU against the GPU drop-in.  sigma = 0 so that the runs can be compared particle by particle; forces
    differ by ~2e-6, so steps (sqrt(eps/|a|)) and with them the time line drift apart at that level."""
    if not refdrv_mod.available(kind):
        pytest.skip(f"oracle/_ref/libsidmref_{kind}.so not built")
    cpu = _run_loop("diag", str(tmp_path / "cpu.npz"))
    gpu = _run_loop(kind, str(tmp_path / "gpu.npz"))
    np.testing.assert_allclose(gpu["maxpred0"], cpu["maxpred0"], rtol=1e-5)          # first steps from the start-up forces
    np.testing.assert_allclose(gpu["time"], cpu["time"], rtol=2e-5)                  # the same sequence of system times
    assert cpu["nactive"].min() < N // 20 and cpu["nactive"].max() > N // 4          # small and large active sets occur
    # the same particles are advanced in (nearly) every iteration: near-ties in MaxPredTime may swap between iterations
    assert np.abs(gpu["nactive"] - cpu["nactive"]).sum() <= 0.01 * cpu["nactive"].sum()
    same = gpu["curtime"] == cpu["curtime"]
    assert same.mean() > 0.98
    np.testing.assert_allclose(gpu["pos"][same], cpu["pos"][same], rtol=0, atol=2e-4)
    np.testing.assert_allclose(gpu["vel"][same], cpu["vel"][same], rtol=0, atol=2e-3)
    assert (gpu["ngb"][same] == cpu["ngb"][same]).mean() > 0.99
    assert (np.abs(gpu["ngb"] - 30) <= 2).all()                                      # the repair loop kept every count in range


def test_scatterlog_file_through_the_shim(tmp_path, refdrv_mod):
    """-DSCATTERLOG: the fast drop-in appends one struct scatlog per scattering to sct_<snapshot>.<task> like sidm.c:571-601.
    The two runs draw different random numbers (MT19937 / Philox), so the files agree in format and in statistics: every
    record is an elastic kick between two particles within each other's reach, one record per kicked pair, and the number
    of records is the all-CPU run's within Poisson noise."""
    if not refdrv_mod.available("b200f"):
        pytest.skip("oracle/_ref/libsidmref_b200f.so not built")
    runs = {}
    for kind in ("diag", "b200f"):
        out = str(tmp_path / f"{kind}.npz")
        r = subprocess.run([sys.executable, os.path.join(HERE, "dropin_runner.py"), kind, out, str(N), "sct"], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        runs[kind] = np.load(out)
    cpu, gpu = runs["diag"]["log"], runs["b200f"]["log"]
    assert len(cpu) > 100 and abs(len(gpu) - len(cpu)) < 6 * np.sqrt(len(cpu) + len(gpu))
    for log, r in ((cpu, runs["diag"]), (gpu, runs["b200f"])):
        assert (log["id1"] != log["id2"]).all() and log["id1"].min() >= 1 and log["id2"].max() <= N
        d = np.sqrt(((log["x1"].astype(np.float64) - log["x2"]) ** 2).sum(1))
        assert (d < log["h1"]).all()                                              # the partner lies inside the scatterer's sphere
        v_rel0 = np.sqrt(((log["v1"].astype(np.float64) - log["v2"]) ** 2).sum(1))
        v_rel1 = np.sqrt(((log["v1"].astype(np.float64) + log["dv"] - (log["v2"].astype(np.float64) - log["dv"])) ** 2).sum(1))
        np.testing.assert_allclose(v_rel1, v_rel0, rtol=2e-5)                     # elastic: equal masses keep |v1 - v2|
        kicked = (r["dvel"] != 0).any(axis=1).sum()
        assert len(log) <= kicked <= 2 * len(log)
