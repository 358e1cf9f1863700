"""BASELINE config C2: isolated NFW halo N=1e6, sigma/m = 1 cm^2/g, one B200 - tree + SIDM step, with the tree
accelerations checked against direct summation on a 4096-target subsample."""
import sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sidm-nbody_b200"))
import torch
from sidm_b200 import HotPath, ic

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
pos, vel, mass, ids = ic.nfw(n, seed=2)
hp = HotPath(n, CrossSectionInternal=ic.cross_section_internal(1.0), Seed=55)
hp.set_particles(pos, vel, mass, ids)
hp.predict_collisionless_only(0.0)
hp.force_treebuild()
hp.setup_smoothinglengths_sidm(30)
vmax = hp.getvmax()
idx = np.sort(np.random.default_rng(7).choice(n, 4096, replace=False)).astype(np.int32)
direct = hp.force_treeevaluate_direct(idx)
rms = lambda a, b: float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))
acc_bh, cost_bh = hp.force_treeevaluate(idx)                       # OldAcc = 0: BH criterion, theta = 0.5
hp.compute_accelerations(1, time=0.0, vmax=vmax)                   # start-up forces -> OldAcc
acc_rel, cost_rel = hp.force_treeevaluate(idx)                     # relative criterion, alpha = 0.005
res = dict(n=n, tree_vs_direct_bh=rms(acc_bh, direct), tree_vs_direct_relative=rms(acc_rel, direct),
           interactions_bh=float(cost_bh.sum(1).mean()), interactions_relative=float(cost_rel.sum(1).mean()))
t, dt = 0.0, 1e-4
for _ in range(3):
    hp.compute_accelerations(0, time=t + dt / 2, vmax=vmax); hp.advance(time=t + dt / 2); t += dt
torch.cuda.synchronize(); t0 = time.perf_counter(); reps = 10
for _ in range(reps):
    hp.compute_accelerations(0, time=t + dt / 2, vmax=vmax); hp.advance(time=t + dt / 2); t += dt
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / reps * 1e3
c = hp.counters()
res.update(ms_per_step=round(ms, 3), updates_per_s=round(n / ms * 1e3), build_ms=round(c.ms_build, 3), walk_ms=round(c.ms_walk, 3),
           sidm_ms=round(c.ms_sidm, 3), ensure_ms=round(c.ms_ensure, 3))
print(json.dumps(res))
hp.close()
