"""GPU: k_advance (b200_advance) against the unmodified reference's advance() (predict.c:245-345), bit for bit:
Pos, Vel, VelPred, CurrentTime of the active particles, the dVel clear, untouched inactive particles and
n_scat_particles (predict.c:258,268-269).  Tolerance: none - every operand is a float, the arithmetic is double
in the reference's order of operations (the library is compiled with -fmad=false for these expressions)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 20000
MAXPART = 60000      # the reference allocates once per process (allocate.c:168-185): as large as any later test of this session


def _state(seed):
    rng = np.random.default_rng(seed)
    from sidm_b200 import ic
    pos, vel, mass, ids = ic.hernquist(N, seed=seed)
    accel = (rng.standard_normal((N, 3)) * 3.0e3).astype(np.float32)
    dvel = np.zeros((N, 3), np.float32)
    hit = rng.random(N) < 0.07
    dvel[hit] = (rng.standard_normal((int(hit.sum()), 3)) * 40.0).astype(np.float32)
    dvel[rng.random(N) < 0.01, 0] = 0.0                       # kicks whose x component is exactly 0 do not count (predict.c:268)
    curtime = (rng.random(N) * 0.004).astype(np.float32)
    active = np.sort(rng.choice(N, size=N // 3, replace=False)).astype(np.int32)
    return pos, vel, mass, ids, accel, dvel, curtime, active


@pytest.mark.parametrize("comoving", [0, 1])
def test_advance_bit_exact(refdrv_mod, comoving):
    from sidm_b200 import HotPath
    pos, vel, mass, ids, accel, dvel, curtime, active = _state(5 + comoving)
    t = 0.0123 if not comoving else 0.31
    if comoving:
        curtime = (0.3 + curtime).astype(np.float32)
    cos = dict(ComovingIntegrationOn=comoving, Omega0=0.3, OmegaLambda=0.7, Hubble=0.1)
    R = refdrv_mod.Reference("diag")
    R.setup(MAXPART, **cos)
    R.set_particles(pos, vel, mass, ids)
    R.set("ACCEL", accel); R.set("DVEL", dvel); R.set("CURTIME", curtime)
    R.set_time(t)
    R.set_active(active)
    nscat_r = R.advance()
    with HotPath(N, **cos) as hp:
        hp.set_particles(pos, vel, mass, ids, curtime=curtime, accel=accel, dvel=dvel)
        nscat = hp.advance(active=active, time=t, count=True)
        pos0 = hp.peek("pos0", np.float32, (N, 3))
        velh = hp.peek("velh", np.float32, (N, 4))
        ct = hp.peek("curtime", np.float32, (N,))
        velpred, dv = hp.get("VelPred", "dVel")
    assert nscat == nscat_r > 100
    assert np.array_equal(pos0.view(np.uint32), R.get("POS").view(np.uint32))
    assert np.array_equal(velh[:, :3].copy().view(np.uint32), R.get("VEL").view(np.uint32))
    assert np.array_equal(ct.view(np.uint32), R.get("CURTIME").view(np.uint32))
    assert np.array_equal(dv.view(np.uint32), R.get("DVEL").view(np.uint32))
    # VelPred: the reference writes it for the active particles only; set_particles() starts it at Vel on both sides
    assert np.array_equal(velpred[active].view(np.uint32), R.get("VELPRED")[active].view(np.uint32))
    inactive = np.setdiff1d(np.arange(N), active)
    assert np.array_equal(pos0[inactive], pos[inactive]) and np.array_equal(dv[inactive], dvel[inactive])
    assert not dv[active].any()
