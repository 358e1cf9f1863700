"""one launch of the gravity walk on the bench workload, for ncu (scripts/profile_walk_r2.sh).
usage: python scripts/prof_walk.py N pairs(0|1) [minb]"""
import sys
import numpy as np
sys.path.insert(0, "sidm-nbody_b200")
from sidm_b200 import HotPath, ic

N = int(float(sys.argv[1])); pairs = int(sys.argv[2]); minb = int(sys.argv[3]) if len(sys.argv) > 3 else 6
pos, vel, mass, ids = ic.nfw(N, seed=3)
hp = HotPath(N)
hp.set_particles(pos, vel, mass, ids)
hp.predict_collisionless_only(0.0)
hp.set_option("walk_pairs", 0)
hp.force_treebuild()
hp.gravity_tree()                      # BH start-up pass (k_walk) -> OldAcc
hp.set_option("walk_pairs", pairs); hp.set_option("walkp_minb", minb)
hp.force_treebuild()
hp.gravity_tree()                      # the profiled launch: relative criterion
c = hp.counters()
print(f"walk {c.ms_walk:.3f} ms I_n {c.list_nodes / c.num_lists:.1f}")
hp.close()
