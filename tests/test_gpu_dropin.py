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
