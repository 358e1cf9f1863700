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


def test_scatter_replay_three_types(world):
    """sidm() with three types: neighbours and partners come from the particle's own type only (sidm.c:319-330,
    one search tree per type), slots in exported-first order with DomainMin/Max of the particle's type.  The
    reference's own random numbers are replayed: its draw log is split into per-slot uniforms and directions
    with the P_max / total probability the GPU reports per slot (a mismatch derails the split and fails)."""
    R, hp, types = world["R"], world["hp"], world["types"]
    pos, vel, mass, ids = world["pos"], world["vel"], world["mass"], world["ids"]
    SIG, DT = 20.89, 0.02
    R.setup(N, CrossSectionInternal=SIG)
    for t, e in EPS.items():
        R.set_softening(t, e)
    R.init_rand(55)
    R.set_particles(pos, vel, mass, ids)
    R.set("TYPE", types)
    R.treebuild()
    R.setup_smoothinglengths_sidm(30)
    h = R.get("HSML")
    R.all_active(0.0, DT)
    t = R.time
    vmax = R.getvmax()
    R.rng_log_begin(4 * N + 1000)
    R.sidm()
    log = R.rng_log_end()
    dv_ref, ngb_ref = R.get("DVEL"), R.get("NGB")
    nscat = int((np.abs(dv_ref).sum(1) > 0).sum())
    assert nscat >= 10, "fixture too quiet"
    act = np.arange(N, dtype=np.int32)
    hp.set_params(CrossSectionInternal=SIG)
    hp.set_option("cand_cap", 16384)       # pass A runs every slot through pass 2: a search cube in the outskirts can clip the centre
    hp.set_particles(pos, vel, mass, ids, hsml=h, curtime=np.zeros(N, np.float32), dvel=np.zeros((N, 3), np.float32))
    hp.set_field("ptype", types)
    hp.force_treebuild()
    # pass A: any uniforms, only to read P_max and the total probability of every slot
    hp.sidm(active=act, time=t, vmax=vmax, replay_rand=np.full(N, 1e-300), replay_dir=np.zeros((N, 3)))
    sp, pmax, ptot, partner = hp.sidm_debug(N)
    rand = np.zeros(N)
    dirs = np.zeros((N, 3))
    k = 0
    for s in range(N):                      # split the reference's draw log (sidm.c:341, sidm_rand.h:24-37)
        assert k < len(log), f"draw log exhausted at slot {s}"
        u = log[k]
        k += 1
        rand[s] = u
        if pmax[s] < u or not (ptot[s] >= u):
            continue
        while True:
            y1 = 1.0 - 2.0 * log[k]
            y2 = 1.0 - 2.0 * log[k + 1]
            k += 2
            r2 = y1 * y1 + y2 * y2
            if r2 <= 1.0:
                break
        sq = np.sqrt(1.0 - r2)
        dirs[s] = (2.0 * y1 * sq, 2.0 * y2 * sq, 1.0 - 2.0 * r2)
    assert k == len(log), (k, len(log))     # every draw accounted for
    hp.set_particles(hsml=h, dvel=np.zeros((N, 3), np.float32), curtime=np.zeros(N, np.float32))
    hp.sidm(active=act, time=t, vmax=vmax, replay_rand=rand, replay_dir=dirs)
    sp, pmax, ptot, partner = hp.sidm_debug(N)
    hit = partner >= 0
    assert hit.sum() >= 5
    assert np.array_equal(types[sp[hit]], types[partner[hit]]), "a pair across particle types"
    dv, ngb = hp.get("dVel", "NgbVelDisp")
    assert np.array_equal(ngb, ngb_ref), "neighbour counts"
    assert np.array_equal(dv != 0, dv_ref != 0), "who scattered"
    np.testing.assert_allclose(dv, dv_ref, rtol=3e-6, atol=1e-30)
