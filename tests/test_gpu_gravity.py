"""GPU parity: tree build + walk + direct sum through the C ABI against the unmodified
reference (oracle/_ref) on the same seeded inputs.

Tolerances (BASELINE.json north_star): tree geometry bit-exact; accelerations relative rms
< 1e-4 against the reference tree (we require < 2e-6: same interaction sets, float
arithmetic); both also checked against direct summation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 60000


def rel_rms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


@pytest.fixture(scope="module")
def setup(refdrv_mod):
    from sidm_b200 import HotPath, ic
    pos, vel, mass, ids = ic.hernquist(N, seed=1)
    R = refdrv_mod.Reference("diag")
    R.setup(N)
    R.set_particles(pos, vel, mass, ids)
    R.treebuild()
    hp = HotPath(N)
    hp.set_particles(pos, vel, mass, ids)
    nn = hp.force_treebuild()
    yield dict(R=R, hp=hp, pos=pos, vel=vel, mass=mass, ids=ids, nn=nn)
    hp.close()


def _sorted_view(center, length):
    a = np.concatenate([center, length[:, None]], axis=1).astype(np.float32).copy()
    v = a.view([("a", "<u4"), ("b", "<u4"), ("c", "<u4"), ("d", "<u4")]).ravel()
    o = np.argsort(v, order=["a", "b", "c", "d"])
    return v[o], o


def test_tree_geometry_bit_exact(setup):
    R, hp = setup["R"], setup["hp"]
    nd = R.dump_nodes()
    t = hp.get_tree()
    assert len(t["len"]) == len(nd["len"]) == setup["nn"]
    kr, orr = _sorted_view(nd["center"], nd["len"])
    kg, og = _sorted_view(t["center"], t["len"])
    assert np.array_equal(kr, kg), "set of (centre, len) cells differs from the reference"
    assert np.array_equal(nd["count"][orr], t["count"][og])
    assert np.array_equal(nd["mass"][orr], t["mass"][og])
    assert np.array_equal(nd["oc"][orr], t["oc"][og])
    np.testing.assert_allclose(t["bmax2"][og], nd["bmax2"][orr], rtol=3e-7)
    np.testing.assert_allclose(t["s"][og], nd["s"][orr], rtol=0, atol=2e-7 * float(np.abs(nd["center"]).max()))
    scale = (nd["mass"][orr] * nd["len"][orr] ** 2)[:, None]
    Qr = np.concatenate([nd["Q"][orr], nd["P"][orr][:, None]], axis=1)
    assert np.max(np.abs(Qr - t["Q"][og]) / scale) < 1e-6


def test_walk_bh_matches_reference(setup):
    R, hp = setup["R"], setup["hp"]
    idx = np.arange(0, N, 29, dtype=np.int32)
    R.set("OLDACC", np.zeros(N, np.float32))
    acc_r, cost_r = R.force_tree(idx)
    hp.set_particles(oldacc=np.zeros(N, np.float32))
    hp.force_treebuild()
    acc, cost = hp.force_treeevaluate(idx)
    assert rel_rms(acc, acc_r) < 2e-6
    assert (cost == cost_r).all(axis=1).mean() > 0.999     # same interaction lists
    d = hp.force_treeevaluate_direct(idx)
    d_r = R.force_direct(idx)
    assert rel_rms(d, d_r) < 1e-6
    assert rel_rms(acc, d) < 5e-3                           # BH theta=0.5 accuracy itself


def test_walk_relative_criterion_matches_reference(setup):
    R, hp = setup["R"], setup["hp"]
    full = np.arange(N, dtype=np.int32)
    R.set("OLDACC", np.zeros(N, np.float32))
    acc_full, _ = R.force_tree(full, want_cost=False)
    a = acc_full.astype(np.float32)
    oa = np.sqrt((a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1] + a[:, 2] * a[:, 2]).astype(np.float64)).astype(np.float32)
    R.set("OLDACC", oa)
    idx = np.arange(3, N, 17, dtype=np.int32)
    acc_r, cost_r = R.force_tree(idx)
    hp.set_particles(oldacc=oa)
    hp.force_treebuild()
    acc, cost = hp.force_treeevaluate(idx)
    assert rel_rms(acc, acc_r) < 2e-6
    assert (cost == cost_r).all(axis=1).mean() > 0.999
    c = hp.counters()
    assert c.part_interactions == int(cost[:, 0].sum()) and c.node_interactions == int(cost[:, 1].sum())


def test_gravity_tree_epilogue(setup):
    """gravity_tree(): Accel = G*acc, OldAcc = |acc| (gravtree.c:300-324) for all particles."""
    R, hp = setup["R"], setup["hp"]
    R.set("OLDACC", np.zeros(N, np.float32))
    R.all_active(0.0, 0.0)
    R.gravity_tree()
    acc_r, oa_r = R.get("ACCEL"), R.get("OLDACC")
    hp.set_particles(oldacc=np.zeros(N, np.float32))
    hp.predict_collisionless_only(0.0)
    hp.force_treebuild()
    hp.gravity_tree()
    acc, oa = hp.get("Accel", "OldAcc")
    assert rel_rms(acc.astype(np.float64), acc_r.astype(np.float64)) < 2e-6
    np.testing.assert_allclose(oa, oa_r, rtol=2e-5)
    # second call uses the relative criterion with the new OldAcc on both sides
    R.gravity_tree()
    hp.gravity_tree()
    acc2_r = R.get("ACCEL")
    acc2 = hp.get("Accel")
    assert rel_rms(acc2.astype(np.float64), acc2_r.astype(np.float64)) < 1e-4


def test_packed_walk_equals_stream_walk(setup):
    """k_walk_pairs (sibling pairs, f32x2) and k_walk (pre-order stream) take the same decisions: identical interaction
    counts for every target, accelerations equal up to the order of the float partial sums"""
    hp = setup["hp"]
    idx = np.arange(0, N, 7, dtype=np.int32)
    for oa in (np.zeros(N, np.float32), None):               # BH start-up criterion, then the relative criterion
        if oa is not None:
            hp.set_particles(oldacc=oa)
        hp.set_option("walk_pairs", 0)
        hp.force_treebuild()
        a0, c0 = hp.force_treeevaluate(idx)
        hp.set_option("walk_pairs", 1)
        for minb in (4, 6, 8):
            hp.set_option("walkp_minb", minb)
            hp.force_treebuild()
            a1, c1 = hp.force_treeevaluate(idx)
            assert np.array_equal(c0, c1)
            assert rel_rms(a1, a0) < 1e-6
        if oa is not None:
            hp.gravity_tree()                                 # OldAcc for the second round
    hp.set_option("walkp_minb", 6)
    hp.set_option("walk_pairs", 0)
    hp.force_treebuild()
