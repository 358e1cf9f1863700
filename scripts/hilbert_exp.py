"""experiment: does grouping the walk's targets by Peano-Hilbert order instead of the tree's own (Morton-like) order shorten the
node stream a warp has to visit?  Same tree, same per-target decisions; only the assignment of targets to warps changes.
usage (GPU): B200_KEEP_TARGET_ORDER=1 python scripts/hilbert_exp.py [N]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sidm-nbody_b200")); sys.path.insert(0, ROOT)
import bench
from sidm_b200 import HotPath


def hilbert_keys(pos, bits=18):
    lo = pos.min(0); ext = (pos.max(0) - lo).max() * 1.0001
    X = np.minimum(((pos - lo) / ext * (1 << bits)).astype(np.int64), (1 << bits) - 1).T.copy()   # [3, n]
    # Skilling's axes -> transpose
    M = 1 << (bits - 1)
    Q = M
    while Q > 1:
        P = Q - 1
        for i in range(3):
            hit = (X[i] & Q) != 0
            X[0] = np.where(hit, X[0] ^ P, X[0])
            t = np.where(hit, 0, (X[0] ^ X[i]) & P)
            X[0] ^= t; X[i] ^= t
        Q >>= 1
    for i in range(1, 3):
        X[i] ^= X[i - 1]
    t = np.zeros_like(X[0]); Q = M
    while Q > 1:
        t = np.where((X[2] & Q) != 0, t ^ (Q - 1), t)
        Q >>= 1
    for i in range(3):
        X[i] ^= t
    key = np.zeros(X.shape[1], np.int64)
    for b in range(bits - 1, -1, -1):
        for i in range(3):
            key = (key << 1) | ((X[i] >> b) & 1)
    return key


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    cfg = bench.CONFIGS["C3"]
    pos, vel, mass, ids = bench.make_ic(cfg, n)
    hp = HotPath(n, CrossSectionInternal=bench.sigma_internal(cfg), Seed=55, **bench.path_params(cfg, n))
    T0 = bench.t_begin(cfg)
    hp.set_particles(pos, vel, mass, ids, curtime=np.full(n, T0, np.float32))
    hp.predict_collisionless_only(T0); hp.force_treebuild()
    hp.compute_accelerations(1, time=T0, vmax=0.0)
    hp.force_treebuild()
    orders = {"tree order (None)": None}
    orders["tree order (explicit list)"] = hp.peek("sidx", np.int32, (n,))
    hk = hilbert_keys(pos.astype(np.float64))
    orders["Peano-Hilbert order"] = np.argsort(hk, kind="stable").astype(np.int32)
    orders["random order"] = np.random.default_rng(1).permutation(n).astype(np.int32)
    for name, o in orders.items():
        for rep in range(2):
            hp.gravity_tree(active=o, time=T0)
        c = hp.counters()
        print(f"{name:30s} k_walk {c.ms_walk:8.3f} ms   nodes streamed per warp {c.list_nodes / c.num_lists:8.1f}   particles per warp {c.list_parts / c.num_lists:7.1f}", flush=True)
    hp.close()


if __name__ == "__main__":
    main()
