"""BASELINE config C5 ingredients: N=4e6 NFW halo, individual time steps = small active sets against the full
tree.  Times b200_compute_accelerations(0, active) (tree rebuild + walk + SIDM + repair) for a few set sizes."""
import sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sidm-nbody_b200"))
import torch
from sidm_b200 import HotPath, ic

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
pos, vel, mass, ids = ic.nfw(n, seed=5)
sig = ic.cross_section_internal(10.0)
hp = HotPath(n, CrossSectionInternal=sig, Seed=55)
hp.set_particles(pos, vel, mass, ids)
hp.predict_collisionless_only(0.0)
hp.force_treebuild()
hp.setup_smoothinglengths_sidm(30)
vmax = hp.getvmax()
hp.compute_accelerations(1, time=0.0, vmax=vmax)
rng = np.random.default_rng(1)
r = np.sqrt((pos.astype(np.float64) ** 2).sum(1))
order = np.argsort(r)                       # the shortest steps live in the centre
out = {}
for na in (1000, 10000, 100000, n):
    act = None if na == n else np.sort(order[:na]).astype(np.int32)
    t = 1e-5
    for _ in range(2):
        hp.compute_accelerations(0, active=act, time=t, vmax=vmax); t += 1e-5
    torch.cuda.synchronize()
    t0 = time.perf_counter(); reps = 5
    for _ in range(reps):
        hp.compute_accelerations(0, active=act, time=t, vmax=vmax); t += 1e-5
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    c = hp.counters()
    out[na] = dict(ms=round(ms, 3), build_ms=round(c.ms_build, 3), walk_ms=round(c.ms_walk, 3), sidm_ms=round(c.ms_sidm, 3),
                   ensure_ms=round(c.ms_ensure, 3), updates_per_s=round(na / ms * 1e3))
    print(na, out[na], flush=True)
print(json.dumps({"n": n, "active_sets": out}))
hp.close()
