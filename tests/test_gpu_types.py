"""GPU: several collisionless particle types - one octree per type, walked one after the other with
epsilon = max(eps of the tree's type, eps of the target's type) (forcetree.c:90-158, 798-808), neighbour searches
and scatterings inside the particle's own type (sidm.c:319-330) - against the unmodified reference."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = 40000
EPS = {1: 0.3, 2: 0.7, 3: 0.15}


def rel_rms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


@pytest.fixture(scope="module")
def world(refdrv_mod):
    from sidm_b200 import HotPath, ic
    pos, vel, mass, ids = ic.hernquist(N, seed=51)
    rng = np.random.default_rng(3)
    types = rng.choice(np.array([1, 2, 3], np.int32), N, p=[0.5, 0.35, 0.15]).astype(np.int32)
    mass = (mass * np.where(types == 2, 2.0, 1.0)).astype(np.float32)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    R = refdrv_mod.Reference("diag")
    R.setup(N, CrossSectionInternal=0.0)
    for t, e in EPS.items():
        R.set_softening(t, e)
    R.init_rand(55)
    R.set_particles(pos, vel, mass, ids)
    R.set("TYPE", types)
    R.treebuild()
    table = [0.0] * 6
    for t, e in EPS.items():
        table[t] = e
    hp = HotPath(N, CrossSectionInternal=0.0, SofteningTable=table, ReferenceNgbOrder=1)
    hp.set_particles(pos, vel, mass, ids)
    hp.set_field("ptype", types)
    hp.force_treebuild()
    yield dict(R=R, hp=hp, pos=pos, vel=vel, mass=mass, ids=ids, types=types)
    hp.close()
    os.chdir(cwd)


def test_forces_three_types(world):
    R, hp = world["R"], world["hp"]
    idx = np.arange(0, N, 7, dtype=np.int32)
    R.set("OLDACC", np.zeros(N, np.float32))
    hp.set_particles(oldacc=np.zeros(N, np.float32))
    a_r, c_r = R.force_tree(idx)                       # BH, three trees
    a, c = hp.force_treeevaluate(idx)
    assert rel_rms(a, a_r) < 2e-6
    assert (c.sum(1) == c_r.sum(1)).mean() > 0.999
    assert rel_rms(hp.force_treeevaluate_direct(idx), R.force_direct(idx)) < 1e-6
    full = np.arange(N, dtype=np.int32)
    af, _ = R.force_tree(full, want_cost=False)
    a32 = af.astype(np.float32)
    oa = np.sqrt((a32[:, 0] * a32[:, 0] + a32[:, 1] * a32[:, 1] + a32[:, 2] * a32[:, 2]).astype(np.float64)).astype(np.float32)
    R.set("OLDACC", oa)
    hp.set_particles(oldacc=oa)
    a_r, c_r = R.force_tree(idx)                       # relative criterion
    a, c = hp.force_treeevaluate(idx)
    assert rel_rms(a, a_r) < 2e-6
    assert (c.sum(1) == c_r.sum(1)).mean() > 0.999
    np.testing.assert_allclose(hp.force_treeevaluate_potential(idx), R.potential(idx), rtol=3e-6)


def test_smoothing_lengths_and_step_three_types(world):
    R, hp, types = world["R"], world["hp"], world["types"]
    R.setup_smoothinglengths_sidm(30)                  # init.c:431: k-NN and counts inside each particle's own type
    hp.setup_smoothinglengths_sidm(30)
    h_r, n_r = R.get("HSML"), R.get("NGB")
    h, n = hp.get("HsmlVelDisp", "NgbVelDisp")
    assert np.array_equal(n, n_r)
    assert (h == h_r).mean() > 0.999
    np.testing.assert_allclose(h, h_r, rtol=3e-7)
    # a full compute_accelerations(0): gravity over three trees + sidm (sigma = 0: counts only) + repair loop
    rng = np.random.default_rng(4)
    h2 = (h_r * rng.choice(np.array([1, 1, 1, 0.8, 1.3], np.float32), N)).astype(np.float32)
    R.set("HSML", h2)
    R.set("OLDACC", np.zeros(N, np.float32))           # BH criterion on both sides (the first test left OldAcc set)
    R.all_active(0.0, 0.01)
    R.getvmax()
    R.compute_accelerations(0)
    hp.set_particles(hsml=h2, curtime=np.zeros(N, np.float32), oldacc=np.zeros(N, np.float32))
    hp.compute_accelerations(0, time=R.time, vmax=0.0)
    acc, ngb, hh, pp = hp.get("Accel", "NgbVelDisp", "HsmlVelDisp", "PosPred")
    assert np.array_equal(pp, R.get("POSPRED"))
    assert rel_rms(acc.astype(np.float64), R.get("ACCEL").astype(np.float64)) < 1e-4
    assert np.array_equal(ngb, R.get("NGB"))
    assert (hh == R.get("HSML")).mean() > 0.999
