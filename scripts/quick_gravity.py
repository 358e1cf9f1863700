"""ad-hoc: time tree build + walk at a given N on the GPU (development aid)."""
import sys, time
import numpy as np
sys.path.insert(0, "sidm-nbody_b200")
from sidm_b200 import HotPath, ic

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
pos, vel, mass, ids = ic.nfw(N, seed=2)
hp = HotPath(N)
hp.set_particles(pos, vel, mass, ids)
hp.predict_collisionless_only(0.0)
for rep in range(3):
    hp.force_treebuild()
    c = hp.counters()
    print(f"build {c.ms_build:.3f} ms nodes {c.num_nodes} maxlev {c.max_level}")
for rep in range(3):
    hp.gravity_tree()
    c = hp.counters()
    print(f"walk rep{rep} {c.ms_walk:.3f} ms  part/target {c.part_interactions/N:.1f} node/target {c.node_interactions/N:.1f} "
          f"list_nodes/warp {c.list_nodes/(N/32):.0f} list_parts/warp {c.list_parts/(N/32):.0f}  -> {N/c.ms_walk/1e3:.2f} M targets/s")
idx = np.arange(0, N, max(1, N // 2048), dtype=np.int32)
acc, cost = hp.force_treeevaluate(idx)
d = hp.force_treeevaluate_direct(idx)
print("tree vs direct rel rms", float(np.sqrt(((acc - d) ** 2).sum() / (d ** 2).sum())))
