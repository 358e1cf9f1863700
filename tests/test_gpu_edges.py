"""GPU edge cases: ragged / tiny particle numbers, empty and partial active lists (individual
time steps, BASELINE config C5: small active sets against the full tree)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel_rms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


@pytest.mark.parametrize("n", [1, 2, 3, 31, 33, 257, 1000])
def test_tiny_and_ragged_sizes(n):
    import oracle
    from sidm_b200 import HotPath, ic
    pos, vel, mass, ids = ic.hernquist(max(n, 8), seed=23)
    pos, vel, mass, ids = pos[:n].copy(), vel[:n].copy(), mass[:n].copy(), ids[:n].copy()
    O = oracle.Oracle(pos, vel, mass)
    m = O.treebuild()
    idx = np.arange(n, dtype=np.int32)
    acc_o, cost_o = O.force_tree(idx)
    with HotPath(n) as hp:
        hp.set_particles(pos, vel, mass, ids)
        assert hp.force_treebuild() == m
        acc, cost = hp.force_treeevaluate(idx)
        assert np.array_equal(cost, cost_o)
        if n > 1:
            assert rel_rms(acc, acc_o) < 5e-6
        else:
            assert not acc.any()
        hp.gravity_tree()                       # every particle, epilogue included
        d = hp.force_treeevaluate_direct(idx)
        assert rel_rms(d, O.force_direct(idx)) < 1e-6 or n == 1


def test_partial_and_empty_active_lists():
    import oracle
    from sidm_b200 import HotPath, ic
    n = 20000
    sigma = 300.0
    pos, vel, mass, ids = ic.hernquist(n, seed=29)
    rng = np.random.default_rng(3)
    active = rng.choice(n, 1500, replace=False).astype(np.int32)      # list order is NOT index order
    O = oracle.Oracle(pos, vel, mass, sigma=sigma)
    O.treebuild()
    with HotPath(n, CrossSectionInternal=sigma, ReferenceNgbOrder=1) as hp:
        hp.set_particles(pos, vel, mass, ids)
        hp.predict_collisionless_only(0.0)
        hp.force_treebuild()
        # gravity for the active particles only: BH first, then relative with their new OldAcc
        hp.gravity_tree(active=active)
        acc_o, _ = O.force_tree(active)
        a_o, oa_o = O.epilogue(acc_o)
        acc, oa = hp.get("Accel", "OldAcc")
        assert rel_rms(acc[active].astype(np.float64), a_o.astype(np.float64)) < 2e-6
        inactive = np.setdiff1d(np.arange(n), active)
        assert not acc[inactive].any() and not oa[inactive].any()      # untouched
        oldacc = np.zeros(n, np.float32)
        oldacc[active] = oa_o
        hp.set_particles(oldacc=oldacc)
        hp.gravity_tree(active=active)
        acc2_o, _ = O.force_tree(active, oldacc)
        a2_o, _ = O.epilogue(acc2_o)
        assert rel_rms(hp.get("Accel")[active].astype(np.float64), a2_o.astype(np.float64)) < 1e-5
        # empty list: nothing happens, no error
        hp.gravity_tree(active=np.zeros(0, np.int32))
        # SIDM on the partial list, reference random numbers replayed
        hp.setup_smoothinglengths_sidm(30)
        h = hp.get("HsmlVelDisp")
        O.hsml[:] = h
        O.init_rand(55)
        vmax = O.getvmax()
        dt = 0.02
        res = O.sidm(active, np.float32(dt), vmax)
        assert res["sct"][2] > 0
        hp.set_particles(dvel=np.zeros((n, 3), np.float32), curtime=np.zeros(n, np.float32))
        hp.sidm(active=active, time=dt / 2, vmax=vmax, replay_rand=res["rand"], replay_dir=res["dir"])
        sp, pmax, ptot, partner = hp.sidm_debug(len(active))
        assert np.array_equal(sp, res["slot_particle"])
        assert np.array_equal(partner, res["partner"])
        dv, ngb = hp.get("dVel", "NgbVelDisp")
        assert np.array_equal(dv != 0, O.dvel != 0)
        assert np.array_equal(ngb[active], O.ngb[active])
        np.testing.assert_allclose(dv, O.dvel, rtol=3e-6, atol=1e-30)


def test_group_search_equals_per_query_search():
    """the warp-shared neighbour search of all-active passes (k_pass1_group) must return the counts
    of the per-query tree search (k_pass1, the restatement of forcetree.c:2163-2297) bit for bit,
    and the same scatter decisions for the same per-particle random numbers"""
    from sidm_b200 import HotPath, ic
    n = 60000
    pos, vel, mass, ids = ic.hernquist(n, seed=11)
    out = {}
    for mode in (0, 1, 2):                                     # 2: warp-shared search with a tiny cell queue = its overflow fallback
        with HotPath(n, CrossSectionInternal=208.9, Seed=9) as hp:
            hp.set_option("group_search", min(mode, 1))
            hp.set_option("queue_cap", 6 if mode == 2 else 320)
            hp.set_particles(pos, vel, mass, ids)
            hp.predict_collisionless_only(0.0)
            hp.force_treebuild()
            hp.setup_smoothinglengths_sidm(30)
            h0, ngb0 = hp.get("HsmlVelDisp", "NgbVelDisp")
            vmax = hp.getvmax()
            hp.set_particles(curtime=np.zeros(n, np.float32))
            hp.sidm(time=0.01, vmax=vmax)                      # all particles, Philox keyed per particle
            ngb1, dvel = hp.get("NgbVelDisp", "dVel")
            c = hp.counters()
            out[mode] = (h0, ngb0, ngb1, dvel, c.sct_pass1, c.sct_scattered)
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2], out[1][2])
    assert np.array_equal(out[0][3], out[1][3])
    assert out[0][4] == out[1][4] and out[0][5] == out[1][5]
    assert out[0][5] > 0
    for k in range(6):
        assert np.array_equal(out[1][k], out[2][k]), f"overflow fallback differs in item {k}"
