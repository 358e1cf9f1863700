"""CPU: the per-index bodies of the CUDA tree build and the per-lane walk arithmetic
(csrc/build_logic.h, csrc/tree_logic.h), run sequentially on the host by tests/hostcheck,
against the golden vectors from the unmodified reference.  This checks the construction
logic the GPU executes (key-prefix octree, pre-order ids, moment shifts, next[] ranks)
where no GPU is available; the GPU run itself is checked by the -m gpu tests."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "hernquist3k.npz")


@pytest.fixture(scope="module")
def hc():
    sys.path.insert(0, os.path.join(HERE, "hostcheck"))
    import build
    return C.CDLL(build.build())


def _sorted_view(center, length):
    a = np.concatenate([center, length[:, None]], axis=1).astype(np.float32).copy()
    v = a.view([("a", "<u4"), ("b", "<u4"), ("c", "<u4"), ("d", "<u4")]).ravel()
    o = np.argsort(v, order=["a", "b", "c", "d"])
    return v[o], o


def test_key_tree_equals_insertion_tree(hc):
    g = np.load(GOLD)
    pos, mass = np.ascontiguousarray(g["pos"]), np.ascontiguousarray(g["mass"])
    n = len(mass)
    assert hc.hc_build(n, pos.ctypes, mass.ctypes, 1) == 0
    m = hc.hc_num_nodes()
    assert m == len(g["node_len"])
    d = dict(center=np.empty((m, 3), np.float32), len=np.empty(m, np.float32), mass=np.empty(m, np.float32),
             s=np.empty((m, 3), np.float32), Q=np.empty((m, 7), np.float32), oc=np.empty(m, np.float32),
             bmax2=np.empty(m, np.float32), count=np.empty(m, np.int32), level=np.empty(m, np.int32),
             skip=np.empty(m, np.int32), parent=np.empty(m, np.int32), minidx=np.empty(m, np.int32))
    hc.hc_get_tree(*[d[k].ctypes for k in ("center", "len", "mass", "s", "Q", "oc", "bmax2", "count", "level", "skip",
                                           "parent", "minidx")])
    kr, orr = _sorted_view(g["node_center"], g["node_len"])
    kh, oh = _sorted_view(d["center"], d["len"])
    assert np.array_equal(kr, kh)
    for k in ("count", "mass", "oc"):
        assert np.array_equal(g["node_" + k][orr], d[k][oh]), k
    np.testing.assert_allclose(d["bmax2"][oh], g["node_bmax2"][orr], rtol=3e-7)
    scale = (g["node_mass"][orr] * g["node_len"][orr] ** 2)[:, None]
    assert np.max(np.abs(g["node_Q"][orr] - d["Q"][oh]) / scale) < 1e-6
    # pre-order invariants the walk relies on
    ids = np.arange(m)
    assert np.all(d["skip"] > ids) and d["skip"][0] == m
    assert np.all(d["parent"][1:] < ids[1:]) and np.all(d["parent"][1:] >= 0)
    # next[] chain rank
    sidx, lo, lr = (np.empty(n, np.int32) for _ in range(3))
    hc.hc_get_orders(sidx.ctypes, lo.ctypes, lr.ctypes)
    assert np.array_equal(np.argsort(lr), g["chain"])
    # walk arithmetic: same interaction lists as the reference, float-level agreement
    idx = np.ascontiguousarray(g["idx"])
    acc = np.empty((len(idx), 3))
    cost = np.empty((len(idx), 2), np.int32)
    zero = np.zeros(n, np.float32)
    hc.hc_walk(len(idx), idx.ctypes, zero.ctypes, 1, C.c_float(0.5), C.c_float(0.005), C.c_float(0.3), acc.ctypes, cost.ctypes)
    assert np.array_equal(cost, g["cost_bh"])
    assert np.sqrt(((acc - g["acc_bh"]) ** 2).sum() / (g["acc_bh"] ** 2).sum()) < 1e-6
    oa = np.ascontiguousarray(g["oldacc"])
    hc.hc_walk(len(idx), idx.ctypes, oa.ctypes, 1, C.c_float(0.5), C.c_float(0.005), C.c_float(0.3), acc.ctypes, cost.ctypes)
    assert (cost == g["cost_rel"]).all(axis=1).mean() > 0.995
    assert np.sqrt(((acc - g["acc_rel"]) ** 2).sum() / (g["acc_rel"] ** 2).sum()) < 1e-5


def _tree(hc):
    m = hc.hc_num_nodes()
    d = dict(center=np.empty((m, 3), np.float32), len=np.empty(m, np.float32), mass=np.empty(m, np.float32),
             s=np.empty((m, 3), np.float32), Q=np.empty((m, 7), np.float32), oc=np.empty(m, np.float32),
             bmax2=np.empty(m, np.float32), count=np.empty(m, np.int32), level=np.empty(m, np.int32),
             skip=np.empty(m, np.int32), parent=np.empty(m, np.int32), minidx=np.empty(m, np.int32))
    hc.hc_get_tree(*[d[k].ctypes for k in ("center", "len", "mass", "s", "Q", "oc", "bmax2", "count", "level", "skip",
                                           "parent", "minidx")])
    return d


def test_forest_of_types_equals_one_tree_per_type(hc):
    """several particle types (forcetree.c:90-158: one tree per type): the forest the build lays out in one node array
    - root cell per type, keys along the particle's own tree, trees one after the other - is, tree by tree, exactly the
    single tree of that type's particles alone (which test_key_tree_equals_insertion_tree pins on the reference)"""
    g = np.load(os.path.join(HERE, "golden", "global3k.npz"))
    pos, mass, types = np.ascontiguousarray(g["pospred"]), np.ascontiguousarray(g["mass"]), np.ascontiguousarray(g["types"])
    n = len(mass)
    assert hc.hc_build_types(n, pos.ctypes, mass.ctypes, types.ctypes, 1) == 0
    forest = _tree(hc)
    sidx, lo, lr = (np.empty(n, np.int32) for _ in range(3))
    hc.hc_get_orders(sidx.ctypes, lo.ctypes, lr.ctypes)
    roots = np.nonzero(forest["parent"] < 0)[0]
    present = [t for t in range(6) if (types == t).any()]
    assert len(roots) == len(present) == 3 and roots[0] == 0
    bounds = list(roots) + [len(forest["len"])]
    assert np.array_equal(types[sidx], np.sort(types))                     # tree after tree in the sorted order
    at = 0
    for k, t in enumerate(present):
        sel = np.nonzero(types == t)[0]
        p_t, m_t = np.ascontiguousarray(pos[sel]), np.ascontiguousarray(mass[sel])
        assert hc.hc_build(len(sel), p_t.ctypes, m_t.ctypes, 1) == 0
        one = _tree(hc)
        s1, l1, r1 = (np.empty(len(sel), np.int32) for _ in range(3))
        hc.hc_get_orders(s1.ctypes, l1.ctypes, r1.ctypes)
        a, b = bounds[k], bounds[k + 1]
        assert b - a == len(one["len"]), t
        for f in ("center", "len", "mass", "s", "Q", "oc", "bmax2", "count", "level"):
            assert np.array_equal(forest[f][a:b], one[f]), (t, f)
        assert np.array_equal(forest["skip"][a:b] - a, one["skip"])       # the root's skip is the next tree's root
        assert np.array_equal(np.where(forest["parent"][a:b] < 0, -1, forest["parent"][a:b] - a), one["parent"])
        assert np.array_equal(sidx[at:at + len(sel)], sel[s1])             # the same particle order inside the tree
        assert np.array_equal(np.argsort(lr[sel], kind="stable"), np.argsort(r1, kind="stable"))   # and the same next[] chain
        at += len(sel)


def test_forest_walk_matches_reference_golden(hc):
    """the walk's lane arithmetic over three trees (k_walk<PER, MULTI>: every tree in turn, h = 2.8 max(eps_tree,
    eps_target)) against the reference's three-type golden vectors: identical interaction counts, float-level forces"""
    g = np.load(os.path.join(HERE, "golden", "global3k.npz"))
    t = np.load(os.path.join(HERE, "golden", "types3k.npz"))
    pos, mass, types = np.ascontiguousarray(g["pospred"]), np.ascontiguousarray(g["mass"]), np.ascontiguousarray(g["types"])
    n = len(mass)
    assert hc.hc_build_types(n, pos.ctypes, mass.ctypes, types.ctypes, 0) == 0
    idx = np.ascontiguousarray(t["idx"])
    eps = np.ascontiguousarray(g["eps"], np.float32)
    acc = np.empty((len(idx), 3))
    cost = np.empty((len(idx), 2), np.int32)
    zero = np.zeros(n, np.float32)
    hc.hc_walk_types(len(idx), idx.ctypes, zero.ctypes, types.ctypes, eps.ctypes, 1, C.c_float(0.5), C.c_float(0.005), acc.ctypes, cost.ctypes)
    assert (cost == t["cost_bh"]).all(axis=1).mean() > 0.995
    assert np.sqrt(((acc - t["acc_bh"]) ** 2).sum() / (t["acc_bh"] ** 2).sum()) < 1e-6
    oa = np.ascontiguousarray(g["oldacc"])
    hc.hc_walk_types(len(idx), idx.ctypes, oa.ctypes, types.ctypes, eps.ctypes, 1, C.c_float(0.5), C.c_float(0.005), acc.ctypes, cost.ctypes)
    assert (cost == t["cost_rel"]).all(axis=1).mean() > 0.995
    assert np.sqrt(((acc - t["acc_rel"]) ** 2).sum() / (t["acc_rel"] ** 2).sum()) < 1e-5


def test_periodic_cubes_clear_of_the_faces_see_no_image(hc):
    """sidm.cu box_interior(): a search cube classified `clear of the faces` must give the SAME neighbour set with the SAME float
    r^2 through the plain distance test as through the wrapped one (ngb_periodic(), forcetree.c:1999-2006, 2195-2206) - also when
    particles stick out of the box, as they do between two do_box_wrapping() calls (run.c:135).  Cubes near a face must not be
    classified clear whenever an image would make a difference."""
    rng = np.random.default_rng(3)
    hc.hc_cube_clear_of_faces.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
    hc.hc_dist2_both.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    n_clear = n_near = n_differs_near = 0
    for box, stick in ((100.0, 0.0), (100.0, 1.5), (7.25, 0.3), (50000.0, 900.0)):
        n = 4000
        pos = (rng.random((n, 3)) * box).astype(np.float32)
        out = rng.random(n) < 0.05                               # 5 % of the particles have drifted out of the box
        pos[out, rng.integers(0, 3, out.sum())] = np.where(rng.random(out.sum()) < 0.5, -rng.random(out.sum()) * stick,
                                                           box + rng.random(out.sum()) * stick).astype(np.float32)
        dom = np.concatenate([pos.min(0), pos.max(0)]).astype(np.float32)
        r2w, r2p = np.empty(n, np.float32), np.empty(n, np.float32)
        for _ in range(400):
            h = np.float32(box * (0.01 + 0.08 * rng.random()))
            # half of the centres near a face, where the classification matters
            c = (rng.random(3) * box).astype(np.float32)
            if rng.random() < 0.5:
                k = rng.integers(0, 3)
                c[k] = np.float32(rng.choice([0.0, box]) + (rng.random() - 0.5) * 4 * (float(h) + stick))
                c[k] = min(max(c[k], np.float32(-stick)), np.float32(box + stick))
            cube = np.concatenate([c - h, c + h]).astype(np.float32)
            clear = hc.hc_cube_clear_of_faces(dom.ctypes.data, box, cube.ctypes.data)
            hc.hc_dist2_both(n, pos.ctypes.data, c.ctypes.data, box, r2w.ctypes.data, r2p.ctypes.data)
            h2 = np.float32(h) * np.float32(h)
            in_w, in_p = r2w < h2, r2p < h2
            if clear:
                n_clear += 1
                assert np.array_equal(in_w, in_p)
                assert np.array_equal(r2w[in_w].view(np.uint32), r2p[in_w].view(np.uint32))
                # and inside the whole search cube no coordinate difference was wrapped (candidates are tested identically)
                inside = np.all((pos >= cube[:3]) & (pos <= cube[3:]), axis=1)
                assert np.array_equal(r2w[inside].view(np.uint32), r2p[inside].view(np.uint32))
            else:
                n_near += 1
                n_differs_near += int(not np.array_equal(in_w, in_p))
    assert n_clear > 400 and n_near > 200
    assert n_differs_near > 20, "the fixture never exercised a periodic image"


def _tree(hc):
    m = hc.hc_num_nodes()
    d = dict(center=np.empty((m, 3), np.float32), len=np.empty(m, np.float32), mass=np.empty(m, np.float32),
             s=np.empty((m, 3), np.float32), Q=np.empty((m, 7), np.float32), oc=np.empty(m, np.float32),
             bmax2=np.empty(m, np.float32), count=np.empty(m, np.int32), level=np.empty(m, np.int32),
             skip=np.empty(m, np.int32), parent=np.empty(m, np.int32), minidx=np.empty(m, np.int32))
    hc.hc_get_tree(*[d[k].ctypes for k in ("center", "len", "mass", "s", "Q", "oc", "bmax2", "count", "level", "skip",
                                           "parent", "minidx")])
    r = dict(len2=np.empty(m, np.float32), first=np.empty(m, np.int32), end=np.empty(m, np.int32), ext=np.empty(m, np.float32))
    hc.hc_get_refit(*[r[k].ctypes for k in ("len2", "first", "end", "ext")])
    d.update(r)
    return d


def test_refit_keeps_the_topology_and_recomputes_every_cell(hc):
    """tree reuse (csrc/tree_build.cu tree_refit_impl, build_logic.h b5_body with BuildView::next): after the particles have
    moved, a refit leaves the cells and their membership as built, recomputes mass / centre of mass / quadrupole of every cell
    from the members' NEW positions, and grows the cell size used by the opening tests so that it covers every member - the
    property that lets a target always open the cell it is filed under.  Nothing moved: bit-identical node records."""
    g = np.load(GOLD)
    pos, mass = np.ascontiguousarray(g["pos"]), np.ascontiguousarray(g["mass"])
    n = len(mass)
    hc.hc_refit.restype = C.c_float
    assert hc.hc_build(n, pos.ctypes, mass.ctypes, 0) == 0
    t0 = _tree(hc)
    assert np.array_equal(t0["len2"], (t0["len"] * t0["len"]).astype(np.float32))
    # nothing moved
    assert hc.hc_refit(pos.ctypes) == 0.0
    t1 = _tree(hc)
    for k in ("mass", "s", "Q", "oc", "bmax2", "len2", "skip", "first", "end"):
        assert np.array_equal(t0[k].view(np.uint32) if t0[k].dtype == np.float32 else t0[k], t1[k].view(np.uint32) if t1[k].dtype == np.float32 else t1[k]), k
    # particles moved by up to 5 % of their radius
    rng = np.random.default_rng(8)
    r = np.sqrt((pos.astype(np.float64) ** 2).sum(1))
    new = (pos + (rng.random((n, 3)) - 0.5) * 0.1 * r[:, None]).astype(np.float32)
    pad = hc.hc_refit(new.ctypes)
    assert pad == np.abs(new - pos).max()
    t2 = _tree(hc)
    for k in ("skip", "first", "end", "count", "parent", "center", "len"):
        assert np.array_equal(t0[k], t2[k]), k                  # topology and geometry as built
    assert np.array_equal(t0["mass"].view(np.uint32), t2["mass"].view(np.uint32))
    lp, lo = np.empty((n, 3), np.float32), np.empty(n, np.int32)
    hc.hc_get_leaves(lp.ctypes, lo.ctypes)
    assert np.array_equal(lp, new[lo])
    grown = 0
    for id in rng.choice(len(t2["len"]), 600, replace=False):
        mem = lo[t2["first"][id]:t2["end"][id]]
        assert len(mem) == t2["count"][id]
        x, w = new[mem].astype(np.float64), mass[mem].astype(np.float64)
        com = (x * w[:, None]).sum(0) / w.sum()
        L = float(t2["len"][id])
        assert np.abs(t2["s"][id] - com).max() < 2e-6 * (np.abs(com).max() + L)
        d = x - com
        q = np.array([(w * d[:, 0] ** 2).sum(), (w * d[:, 1] ** 2).sum(), (w * d[:, 2] ** 2).sum(), (w * d[:, 0] * d[:, 1]).sum(),
                      (w * d[:, 0] * d[:, 2]).sum(), (w * d[:, 1] * d[:, 2]).sum(), (w * (d ** 2).sum(1)).sum()])
        assert np.abs(t2["Q"][id] - q).max() < 1e-5 * (w.sum() * max(L, float(np.abs(d).max())) ** 2)
        ext = np.abs(x - t2["center"][id].astype(np.float64)).max()      # largest coordinate distance of a member from the cell centre
        assert t2["ext"][id] >= ext * (1 - 1e-6)
        side = np.sqrt(float(t2["len2"][id]))
        assert side >= L * (1 - 1e-6) and side / 2 >= ext * (1 - 1e-6)    # the cell of the opening tests covers its members
        if ext > L / 2:
            grown += 1
            assert abs(side - 2 * t2["ext"][id]) < 1e-5 * side
        else:
            assert t2["len2"][id] == t0["len2"][id] or side > L
    assert grown > 20, "the fixture never moved a particle out of its cell"
