"""A/B of the gravity walk kernels on the bench workload (NFW, relative criterion with OldAcc from a start-up pass):
the stack-free pre-order walk (k_walk) against the packed sibling-pair walk (k_walk_pairs) at its occupancy variants.
Prints kernel ms (CUDA events around the launch), list lengths and the agreement of the two results.
usage: python scripts/walk_ab.py [N]"""
import sys
import numpy as np
sys.path.insert(0, "sidm-nbody_b200")
from sidm_b200 import HotPath, ic

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
pos, vel, mass, ids = ic.nfw(N, seed=3)
hp = HotPath(N, CrossSectionInternal=ic.cross_section_internal(1.0))
hp.set_particles(pos, vel, mass, ids)
hp.predict_collisionless_only(0.0)
hp.set_option("walk_pairs", 0)
hp.force_treebuild()
hp.gravity_tree()                      # BH start-up pass -> OldAcc
oa = hp.get("OldAcc")
idx = np.arange(0, N, max(1, N // 65536), dtype=np.int32)


def run(tag, reps=3):
    ms = []
    for _ in range(reps):
        hp.set_particles(oldacc=oa)    # every variant walks with the same OldAcc
        hp.force_treebuild()
        b = hp.counters().ms_build
        hp.gravity_tree()
        c = hp.counters()
        ms.append(c.ms_walk)
    print(f"{tag:28s} walk {min(ms):7.3f} ms (runs {' '.join(f'{m:.2f}' for m in ms)})  build {b:.3f} ms  node/target {c.node_interactions / N:.1f} part/target {c.part_interactions / N:.1f} "
          f"I_n {c.list_nodes / c.num_lists:.1f} I_p {c.list_parts / c.num_lists:.1f}", flush=True)
    hp.set_particles(oldacc=oa)
    hp.force_treebuild()
    return hp.force_treeevaluate(idx)


a0, c0 = run("k_walk (pre-order stream)")
hp.set_option("walk_pairs", 1)
for minb in (4, 6, 8):
    hp.set_option("walkp_minb", minb)
    a1, c1 = run(f"k_walk_pairs minb={minb}")
    rel = float(np.sqrt(((a1 - a0) ** 2).sum() / (a0 ** 2).sum()))
    print(f"    vs k_walk: rel rms {rel:.3e}, identical (particle, node) counts for {(c0 == c1).all(axis=1).mean() * 100:.4f} % of {len(idx)} targets", flush=True)
d = hp.force_treeevaluate_direct(idx[:4096])
print("tree vs direct rel rms (4096 targets)", float(np.sqrt(((a1[:4096] - d) ** 2).sum() / (d ** 2).sum())))
hp.close()
