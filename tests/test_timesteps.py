"""find_timesteps() (timestep.c:17-334): new time steps of the active particles.
CPU: the oracle restatement against the unmodified reference (bit-exact MaxPredTime);
GPU: b200_find_timesteps through the C ABI against the oracle (bit-exact)."""
import os
import tempfile

import numpy as np
import pytest

N = 6000
SIGMA = 208.9
TS = dict(crit=0, eta=0.02, velscale=10.0, probtol=0.2, dyntol=0.05)
T0 = 0.25


def _state():
    """a halo with start-up forces and smoothing lengths from the oracle (CPU only)"""
    import oracle
    from sidm_b200 import ic
    pos, vel, mass, ids = ic.hernquist(N, seed=31)
    O = oracle.Oracle(pos, vel, mass, sigma=SIGMA)
    O.treebuild()
    O.hsml[:] = np.array([np.sqrt(O.ngb_treefind(pos[i], 30)) for i in range(N)], np.float32)
    acc, _ = O.force_tree(np.arange(N, dtype=np.int32))
    accel, _ = O.epilogue(acc)
    rng = np.random.default_rng(5)
    # the state advance() leaves (predict.c:245): All.Time = old MaxPredTime, CurrentTime = All.Time + old dt/2
    maxpred = np.full(N, T0, np.float32)
    curtime = (maxpred + rng.choice(np.array([1e-5, 5e-4, 0.01], np.float32), N)).astype(np.float32)
    return O, pos, vel, mass, ids, accel, curtime, maxpred


@pytest.mark.parametrize("mode,crit", [(2, 0), (1, 0), (1, 1)])
def test_oracle_matches_reference_timesteps(mode, crit, refdrv_mod):
    O, pos, vel, mass, ids, accel, curtime, maxpred = _state()
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        R = refdrv_mod.Reference("diag")
        R.setup(N, CrossSectionInternal=SIGMA)
        R.set_particles(pos, vel, mass, ids)
        R.all_active(0.25, 0.26)
        R.set("ACCEL", accel); R.set("HSML", O.hsml); R.set("CURTIME", curtime); R.set("MAXPRED", maxpred)
        R.set_time(T0)
        vmax = R.getvmax()
        ts = dict(TS, crit=crit)
        R.find_timesteps(mode, **ts)
        ref = R.get("MAXPRED")
    finally:
        os.chdir(cwd)
    got, nclamp = O.find_timesteps(np.arange(N, dtype=np.int32), mode, T0, vmax, accel, curtime, maxpred, **ts)
    assert nclamp == 0
    assert np.array_equal(got, ref)
    assert len(np.unique(got - curtime)) > 100             # genuinely per-particle steps
    if mode != 2:
        dtold = 2 * (curtime.astype(np.float64) + maxpred - 2 * T0)
        assert (np.abs((got.astype(np.float64) - curtime) * 2 - 1.3 * dtold) < 1e-6).any()   # growth limit active somewhere


@pytest.mark.gpu
@pytest.mark.parametrize("mode,crit,clamp", [(2, 0, False), (0, 0, False), (1, 1, False), (1, 0, True)])
def test_gpu_matches_oracle_timesteps(mode, crit, clamp):
    from sidm_b200 import HotPath
    O, pos, vel, mass, ids, accel, curtime, maxpred = _state()
    vmax = O.getvmax()
    ts = dict(TS, crit=crit)
    if clamp:
        ts.update(dtmax=2e-3, dtmin=5e-5)
    rng = np.random.default_rng(9)
    active = np.sort(rng.choice(N, N // 2, replace=False)).astype(np.int32) if mode == 0 else np.arange(N, dtype=np.int32)
    jitter = rng.random(len(active))
    want, nclamp = O.find_timesteps(active, mode, T0, vmax, accel, curtime, maxpred, jitter=jitter, **ts)
    with HotPath(N, CrossSectionInternal=SIGMA) as hp:
        hp.set_particles(pos, vel, mass, ids, curtime=curtime, accel=accel, hsml=O.hsml)
        hp.set_field("maxpred", maxpred)
        out, nc = hp.find_timesteps(mode, active=None if mode != 0 else active, time=T0, vmax=vmax, jitter=jitter, **ts)
        allmp = hp.peek("maxpred", np.float32, (N,))
    assert nc == nclamp and (nclamp > 0) == clamp
    assert np.array_equal(out, want[active])
    assert np.array_equal(allmp, want)                     # inactive particles untouched


def test_oracle_matches_reference_reflect(refdrv_mod):
    """reflect(), reflection.c:7-33"""
    import oracle
    from sidm_b200 import ic
    pos, vel, mass, ids = ic.hernquist(N, seed=33)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        R = refdrv_mod.Reference("diag")
        R.setup(N)
        R.set_particles(pos, vel, mass, ids)
        R.all_active(0.0, 0.01)
        R.reflect(40.0)
        ref = R.get("VEL")
    finally:
        os.chdir(cwd)
    O = oracle.Oracle(pos, vel, mass)
    got, nref = O.reflect(np.arange(N, dtype=np.int32), 40.0, pos, vel)
    assert nref > 50 and np.array_equal(got, ref) and not np.array_equal(got, vel)


@pytest.mark.gpu
def test_gpu_matches_oracle_reflect():
    import oracle
    from sidm_b200 import HotPath, ic
    pos, vel, mass, ids = ic.hernquist(N, seed=33)
    O = oracle.Oracle(pos, vel, mass)
    active = np.arange(0, N, 2, dtype=np.int32)
    want, nref = O.reflect(active, 40.0, pos, vel)
    with HotPath(N) as hp:
        hp.set_particles(pos, vel, mass, ids)
        n = hp.reflect(40.0, active=active)
        got = hp.peek("velh", np.float32, (N, 4))[:, :3]
    assert n == nref > 20
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_comoving_timesteps_match_reference(refdrv_mod):
    """the comoving branch of find_timesteps() (timestep.c:45-93: hubble_a, S(a), a^-2 cross-section, comoving
    G*rho limit) directly against the -DPERIODIC reference build (fresh process: one configuration per library)"""
    if not refdrv_mod.available("periodic"):
        pytest.skip("oracle/_ref/libsidmref_per.so not built")
    import subprocess
    import sys
    code = r'''
import sys, os, tempfile, numpy as np
sys.path.insert(0, "oracle"); sys.path.insert(0, "sidm-nbody_b200")
import refdrv
from sidm_b200 import HotPath, ic
BOX, EPS, A, SIG = 100.0, 0.5, 0.25, 1.5e3
pos, vel, mass, ids = ic.periodic_box(16, seed=8, box=BOX, vel_sigma=60.0)
n = len(mass)
rng = np.random.default_rng(2)
accel = (rng.normal(size=(n, 3)) * 40.0).astype(np.float32)
hsml = (BOX / 16 * rng.uniform(1.0, 2.0, n)).astype(np.float32)
maxpred = np.full(n, A, np.float32)
curtime = (maxpred + rng.choice(np.array([1e-5, 2e-4, 3e-3], np.float32), n)).astype(np.float32)
ts = dict(crit=0, eta=0.02, velscale=10.0, probtol=0.2, dyntol=0.05)
root = os.getcwd(); os.chdir(tempfile.mkdtemp())
R = refdrv.Reference("periodic")
cos = dict(ComovingIntegrationOn=1, Omega0=0.3, OmegaLambda=0.7, Hubble=0.1)
R.setup(n, BoxSize=BOX, SofteningHalo=EPS, Time=A, CrossSectionInternal=SIG, **cos)
R.set_particles(pos, vel, mass, ids)
R.all_active(A, A * 1.01)
R.set("ACCEL", accel); R.set("HSML", hsml); R.set("CURTIME", curtime); R.set("MAXPRED", maxpred)
R.set_time(A)
vmax = R.getvmax()
R.find_timesteps(1, **ts)
ref = R.get("MAXPRED")
os.chdir(root)
with HotPath(n, BoxSize=BOX, PeriodicBoundariesOn=1, SofteningHalo=EPS, CrossSectionInternal=SIG, **cos) as hp:
    hp.set_particles(pos, vel, mass, ids, curtime=curtime, accel=accel, hsml=hsml)
    hp.set_field("maxpred", maxpred)
    out, nc = hp.find_timesteps(1, time=A, vmax=vmax, **ts)
assert nc == 0 and len(np.unique(out - curtime)) > 100
assert np.array_equal(out, ref), float(np.abs(out.astype(np.float64) - ref).max())
print("comoving timesteps bit-exact")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-2500:]
