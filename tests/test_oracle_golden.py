"""CPU: the oracle restatement (oracle/*.c) against the committed golden vectors that
tests/golden/make_golden.py produced by running the unmodified reference.  Bit-exact where
the arithmetic is restated operation for operation (tree, forces, neighbour lists, RNG
stream, scattered pairs)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "hernquist3k.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


@pytest.fixture(scope="module")
def orc(gold):
    import oracle
    O = oracle.Oracle(gold["pos"], gold["vel"], gold["mass"], sigma=float(gold["sigma"]))
    O.treebuild()
    return O


def test_tree_bit_exact(gold, orc):
    d = orc.dump()
    assert orc.random_subnodes() == 0
    for k in ("center", "len", "mass", "s", "Q", "oc", "bmax2", "count"):
        assert np.array_equal(d[k], gold["node_" + k]), k
    assert np.array_equal(orc.chain(), gold["chain"])


def test_forces_bit_exact(gold, orc):
    idx = gold["idx"]
    acc, cost = orc.force_tree(idx)                       # OldAcc = 0 -> BH criterion
    assert np.array_equal(acc, gold["acc_bh"]) and np.array_equal(cost, gold["cost_bh"])
    acc, cost = orc.force_tree(idx, gold["oldacc"])       # relative criterion
    assert np.array_equal(acc, gold["acc_rel"]) and np.array_equal(cost, gold["cost_rel"])
    assert np.array_equal(orc.force_direct(idx), gold["direct"])


def test_gravity_tree_epilogue(gold, orc):
    n = len(gold["mass"])
    full = np.arange(n, dtype=np.int32)
    acc, _ = orc.force_tree(full)
    a, oa = orc.epilogue(acc)
    assert np.array_equal(a, gold["g1_acc"]) and np.array_equal(oa, gold["g1_old"])
    acc2, _ = orc.force_tree(full, oa)
    a2, oa2 = orc.epilogue(acc2)
    assert np.array_equal(a2, gold["g2_acc"]) and np.array_equal(oa2, gold["g2_old"])


def test_neighbours_bit_exact(gold, orc):
    idx, pos, h = gold["idx"], gold["pos"], gold["hsml"]
    knn = np.array([orc.ngb_treefind(pos[i], 30) for i in idx], np.float32)
    assert np.array_equal(knn, gold["knn"])
    for k, i in enumerate(idx):
        lst, _ = orc.ngb_variable(pos[i], h[i])
        ref = gold["nlist"][k]
        ref = ref[ref >= 0]
        assert np.array_equal(lst, ref)


def test_mt19937_known_answer():
    """sidm_rand.c:29-36: after seeding with 55 and 1e6 warm-up draws the reference prints
    0.565737; the next draw is the first one sidm() uses."""
    import oracle
    L = oracle.lib()
    r = L.orng_new(55, 0)
    for _ in range(1000000):
        L.orng_uniform(r)
    v = L.orng_uniform(r)
    L.orng_free(r)
    assert abs(v - 0.5657367655) < 1e-9


def test_sidm_pass_reproduces_reference(gold, orc):
    n = len(gold["mass"])
    orc.hsml[:] = gold["hsml"]
    orc.dvel[:] = 0
    orc.init_rand(55)
    dt32 = np.float32(2 * (float(gold["t_sidm"]) - 0.0))
    res = orc.sidm(np.arange(n, dtype=np.int32), dt32, float(gold["vmax"]))
    assert orc.rng_count() == len(gold["draws"])
    _check_stream(res, gold["draws"])
    assert np.array_equal(orc.dvel, gold["dvel1"]) and np.array_equal(orc.ngb, gold["ngb1"])
    assert np.array_equal(res["log_i"] + 1, gold["log_id1"]) and np.array_equal(res["log_j"] + 1, gold["log_id2"])
    assert np.array_equal(res["log_dv"], gold["log_dv"])


def _check_stream(res, draws):
    """walk the logged MT19937 stream the way sidm() consumes it: one uniform per buffer slot
    (sidm.c:341) and, after a hit, Marsaglia pairs until one is accepted (sidm_rand.h:27-31)."""
    pos = 0
    for s in range(len(res["rand"])):
        assert res["rand"][s] == draws[pos]
        pos += 1
        if res["partner"][s] >= 0:
            while True:
                y1, y2 = 1.0 - 2.0 * draws[pos], 1.0 - 2.0 * draws[pos + 1]
                pos += 2
                r2 = y1 * y1 + y2 * y2
                if r2 <= 1.0:
                    break
            sq = np.sqrt(1.0 - r2)
            assert np.allclose(res["dir"][s], [2 * y1 * sq, 2 * y2 * sq, 1 - 2 * r2], rtol=0, atol=1e-15)
    assert pos == len(draws)


def test_global_quantities_golden():
    """compute_global_quantities_of_system() (global.c:18-135) with three particle types: the 102 doubles of the
    reference's SysState, bit for bit (fixture: make_golden.py global)"""
    import oracle
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "global3k.npz")))
    got = oracle.global_quantities(g["pospred"], g["velpred"], g["mass"], g["pot"], g["types"])
    assert np.array_equal(got, g["sys"])
    # virial sanity of the fixture itself: a Hernquist halo in equilibrium, 2T + W ~ 0
    assert abs(2 * g["sys"][1] + g["sys"][2]) < 0.15 * abs(g["sys"][2])


def test_snapshot_golden():
    """savepositions() (io.c:16-590): the oracle's format-1 writer reproduces the reference's snapshot file of the
    three-type fixture (header bytes and the SHA-256 of the whole file)"""
    import hashlib
    import oracle
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "global3k.npz")))
    got = oracle.snapshot_bytes(g["pospred"], g["velpred"], g["ids"], g["mass"], g["types"], time=float(g["snap_time"]),
                                mass_table=g["snap_mass_table"], hubble_param=0.7, omega0=float(g["snap_omega0"]))
    assert len(got) == int(g["snap_len"])
    assert got[:264] == g["snap_head"].tobytes()
    assert hashlib.sha256(got).hexdigest() == str(g["snap_sha256"])


def test_snapshot_writer_reader_edge_cases(tmp_path):
    """format-1 round trips of the oracle writer / reader: a single particle, every type with a MassTable entry (no
    mass block at all), type-5 particles left out of the file (io.c:265 loops over types 0..4)"""
    import oracle
    rng = np.random.default_rng(3)
    for n, types, mt in ((1, np.array([1], np.int32), None),
                         (50, rng.choice(np.array([1, 2, 3, 4], np.int32), 50), [0, 0.5, 0.25, 2.0, 1.0, 0]),
                         (64, rng.choice(np.array([1, 3, 5], np.int32), 64), [0, 0, 0, 0.125, 0, 0])):
        pos = rng.standard_normal((n, 3)).astype(np.float32); vel = rng.standard_normal((n, 3)).astype(np.float32)
        mass = rng.random(n).astype(np.float32); ids = np.arange(1, n + 1, dtype=np.int32)
        raw = oracle.snapshot_bytes(pos, vel, ids, mass, types, time=0.5, mass_table=mt)
        p = tmp_path / f"snap_{n}"
        p.write_bytes(raw)
        b = oracle.read_snapshot(str(p))
        order = np.concatenate([np.nonzero(types == t)[0] for t in range(5)])
        assert b["npart"].tolist() == [int((types == t).sum()) for t in range(5)] + [0] and b["time"] == 0.5
        assert np.array_equal(b["pos"], pos[order]) and np.array_equal(b["vel"], vel[order]) and np.array_equal(b["ids"], ids[order])
        table = np.zeros(6) if mt is None else np.asarray(mt, np.float64)
        with_mass = order[table[types[order]] == 0]
        if len(with_mass):
            assert np.array_equal(b["mass"], mass[with_mass])
        else:
            assert b["mass"] is None
        nblocks = 4 + (1 if len(with_mass) else 0)
        assert len(raw) == 256 + 28 * len(order) + 4 * len(with_mass) + 8 * nblocks


def test_forest_of_types_golden():
    """three particle types (one tree per type, epsilon = max(eps_tree, eps_target), forcetree.c:90-158, 798-808): the
    oracle forest against the reference's golden vectors - double accelerations with both criteria, interaction
    counts and raw potentials, bit for bit"""
    import oracle
    here = os.path.dirname(__file__)
    g = dict(np.load(os.path.join(here, "golden", "global3k.npz")))
    t = dict(np.load(os.path.join(here, "golden", "types3k.npz")))
    F = oracle.OracleForest(g["pospred"], g["velpred"], g["mass"], g["types"], g["eps"])
    acc, cost = F.force_tree(t["idx"], None)
    assert np.array_equal(acc, t["acc_bh"]) and np.array_equal(cost, t["cost_bh"])
    acc, cost = F.force_tree(t["idx"], g["oldacc"])
    assert np.array_equal(acc, t["acc_rel"]) and np.array_equal(cost, t["cost_rel"])
    assert np.array_equal(F.potential(t["idx"], g["oldacc"]), t["pot_raw"])
    # neighbour searches stay inside the particle's own type (sidm.c:319-330): the counts the reference ended with
    for i in t["idx"][::3]:
        lst, r2 = F.ngb_variable(int(i), t["hsml"][i])
        assert len(lst) == t["ngb"][i] and (g["types"][lst] == g["types"][i]).all()
