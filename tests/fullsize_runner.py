"""Parity at BASELINE.json sizes, run in a fresh process (the reference allocates once per process, allocate.c:168-185).

  c2: configs[1], NFW N = 1e6: tree node set against the unmodified reference (oracle/_ref), walk of 4096 targets with
      both opening criteria against the reference tree and against direct summation, neighbour counts after
      setup_smoothinglengths_sidm() on a 4096 sample.
  c1: configs[0], Hernquist N = 1e5, sigma/m = 1 cm^2/g: scatter pass with the reference's random numbers replayed
      per slot (oracle restatement, pinned bit-exact on the reference): identical pairs.
Prints one JSON object; tests/test_gpu_fullsize.py asserts on it."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (os.path.join(ROOT, "sidm-nbody_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def rel_rms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def sorted_cells(center, length):
    a = np.concatenate([center, length[:, None]], axis=1).astype(np.float32).copy()
    v = a.view([("a", "<u4"), ("b", "<u4"), ("c", "<u4"), ("d", "<u4")]).ravel()
    o = np.argsort(v, order=["a", "b", "c", "d"])
    return v[o], o


def c2(n):
    import refdrv
    from sidm_b200 import HotPath, ic
    out = {}
    pos, vel, mass, ids = ic.nfw(n, seed=2)
    devnull = os.open(os.devnull, os.O_WRONLY)
    keep = os.dup(1)
    os.dup2(devnull, 1)                                   # the reference prints progress to stdout
    try:
        R = refdrv.Reference("diag")
        R.setup(n)
        R.set_particles(pos, vel, mass, ids)
        t0 = time.time()
        R.treebuild()
        out["ref_build_s"] = time.time() - t0
        nd = R.dump_nodes()
        hp = HotPath(n)
        hp.set_particles(pos, vel, mass, ids)
        nn = hp.force_treebuild()
        t = hp.get_tree()
        out["nodes"] = [int(nn), int(len(nd["len"]))]
        kr, orr = sorted_cells(nd["center"], nd["len"])
        kg, og = sorted_cells(t["center"], t["len"])
        out["cells_equal"] = bool(np.array_equal(kr, kg))
        out["count_equal"] = bool(np.array_equal(nd["count"][orr], t["count"][og]))
        out["mass_equal"] = bool(np.array_equal(nd["mass"][orr], t["mass"][og]))
        out["oc_equal"] = bool(np.array_equal(nd["oc"][orr], t["oc"][og]))
        out["s_equal_frac"] = float((nd["s"][orr] == t["s"][og]).all(axis=1).mean())
        scale = (nd["mass"][orr] * nd["len"][orr] ** 2)[:, None]
        Qr = np.concatenate([nd["Q"][orr], nd["P"][orr][:, None]], axis=1)
        out["q_max_rel"] = float(np.max(np.abs(Qr - t["Q"][og]) / scale))
        out["q_equal_frac"] = float((Qr == t["Q"][og]).all(axis=1).mean())
        out["max_level"] = int(hp.counters().max_level)
        # 4096 random targets, BH criterion (OldAcc = 0), then the relative criterion with OldAcc = |a_BH| of the reference
        rng = np.random.default_rng(7)
        idx = np.sort(rng.choice(n, 4096, replace=False)).astype(np.int32)
        R.set("OLDACC", np.zeros(n, np.float32))
        a_r, c_r = R.force_tree(idx)
        hp.set_particles(oldacc=np.zeros(n, np.float32))
        hp.force_treebuild()
        a_g, c_g = hp.force_treeevaluate(idx)
        d_g = hp.force_treeevaluate_direct(idx)
        d_r = R.force_direct(idx[:256])
        out["bh"] = dict(vs_ref=rel_rms(a_g, a_r), same_lists=float((c_g == c_r).all(axis=1).mean()), vs_direct=rel_rms(a_g, d_g),
                         ref_vs_direct=rel_rms(a_r, d_g), direct_vs_ref_direct=rel_rms(d_g[:256], d_r),
                         interactions=float(c_g.sum(axis=1).mean()))
        a32 = a_r.astype(np.float32)
        oa = np.zeros(n, np.float32)
        oa[idx] = np.sqrt((a32[:, 0] * a32[:, 0] + a32[:, 1] * a32[:, 1] + a32[:, 2] * a32[:, 2]).astype(np.float64)).astype(np.float32)
        R.set("OLDACC", oa)
        a_r2, c_r2 = R.force_tree(idx)
        hp.set_particles(oldacc=oa)
        hp.force_treebuild()
        a_g2, c_g2 = hp.force_treeevaluate(idx)
        out["rel"] = dict(vs_ref=rel_rms(a_g2, a_r2), same_lists=float((c_g2 == c_r2).all(axis=1).mean()), vs_direct=rel_rms(a_g2, d_g),
                          ref_vs_direct=rel_rms(a_r2, d_g), interactions=float(c_g2.sum(axis=1).mean()))
        # smoothing lengths of all particles on the GPU, neighbour counts of a sample against the reference's search
        hp.setup_smoothinglengths_sidm(30)
        h, ngb = hp.get("HsmlVelDisp", "NgbVelDisp")
        out["ngb_range"] = [int(ngb.min()), int(ngb.max())]
        cnt = np.array([len(R.ngb_variable(pos[i], float(h[i]))[0]) for i in idx])
        out["ngb_equal"] = bool(np.array_equal(cnt, ngb[idx]))
        h2 = hp.ngb_treefind(idx[:512], 30)
        ref_h2 = np.array([R.ngb_treefind(pos[i], 30) for i in idx[:512]], np.float32)
        out["knn_equal"] = bool(np.array_equal(h2, ref_h2))
        hp.close()
    finally:
        os.dup2(keep, 1)
    return out


def c1(n):
    import oracle
    from sidm_b200 import HotPath, ic
    out = {}
    sigma = ic.cross_section_internal(1.0)
    dt = 0.2
    pos, vel, mass, ids = ic.hernquist(n, seed=1)
    O = oracle.Oracle(pos, vel, mass, sigma=sigma)
    O.treebuild()
    out["random_subnodes"] = int(O.random_subnodes())
    hp = HotPath(n, CrossSectionInternal=sigma, ReferenceNgbOrder=1)
    hp.set_particles(pos, vel, mass, ids)
    hp.force_treebuild()
    hp.setup_smoothinglengths_sidm(30)
    h = hp.get("HsmlVelDisp")
    O.hsml[:] = h
    O.dvel[:] = 0
    O.init_rand(55)
    vmax = O.getvmax()
    active = np.arange(n, dtype=np.int32)
    res = O.sidm(active, np.float32(dt), vmax)
    out["oracle_sct"] = [int(x) for x in res["sct"]]
    hp.set_particles(hsml=h, dvel=np.zeros((n, 3), np.float32), curtime=np.zeros(n, np.float32))
    hp.force_treebuild()
    hp.sidm(active=active, time=dt / 2, vmax=vmax, replay_rand=res["rand"], replay_dir=res["dir"])
    sp, pmax, ptot, partner = hp.sidm_debug(n)
    c = hp.counters()
    out["gpu_sct"] = [c.sct_ntot, c.sct_pass1, c.sct_scattered, c.sct_rejected]
    out["slot_order_equal"] = bool(np.array_equal(sp, res["slot_particle"]))
    out["pmax_max_rel"] = float(np.max(np.abs(pmax - res["pmax"]) / res["pmax"]))
    out["partners_equal"] = bool(np.array_equal(partner, res["partner"]))
    dv, ngb = hp.get("dVel", "NgbVelDisp")
    out["ngb_equal"] = bool(np.array_equal(ngb, O.ngb))
    out["kicked_equal"] = bool(np.array_equal(dv != 0, O.dvel != 0))
    nz = O.dvel != 0
    out["dv_max_rel"] = float(np.max(np.abs(dv[nz] - O.dvel[nz]) / np.abs(O.dvel[nz]))) if nz.any() else 0.0
    log = hp.scatlog()
    out["log_equal"] = bool(np.array_equal(log["id1"], res["log_i"] + 1) and np.array_equal(log["id2"], res["log_j"] + 1))
    nohit = (res["partner"] < 0) & (res["prob"] > 0)
    out["prob_max_rel"] = float(np.max(np.abs(ptot[nohit] - res["prob"][nohit]) / res["prob"][nohit]))
    hp.close()
    return out


if __name__ == "__main__":
    which, n = sys.argv[1], int(float(sys.argv[2]))
    res = c2(n) if which == "c2" else c1(n)
    print("RESULT " + json.dumps(res))
