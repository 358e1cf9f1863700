"""BASELINE config C4 ingredients at full size: periodic comoving box, ngrid^3 particles (default 256^3 =
16 777 216), Ewald-corrected tree gravity + SIDM (sigma/m = 0.5 cm^2/g in h^-1 Mpc units), one GPU.
Functional run with per-phase timings; parity of this path is tested at small size in tests/test_gpu_periodic.py."""
import sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sidm-nbody_b200"))
import torch
from sidm_b200 import HotPath, ic

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 256
BOX, A = 100.0, 0.1
# total mass so that Omega0 = 0.3: rho_crit = 3 H^2 / (8 pi G), H = 0.1 (km/s/kpc -> here 100 km/s/Mpc in box units), G = 43007.1
G, H, O0 = 43007.1, 0.1, 0.3
mtot = O0 * 3 * H * H / (8 * np.pi * G) * BOX ** 3
pos, vel, mass, ids = ic.periodic_box(ng, seed=4, box=BOX, total_mass=mtot, vel_sigma=30.0)
n = len(mass)
sig = ic.cross_section_internal(0.5, unit_length_cm=3.085678e24)
hp = HotPath(n, BoxSize=BOX, PeriodicBoundariesOn=1, SofteningHalo=BOX / ng / 25, ComovingIntegrationOn=1, Omega0=O0, OmegaLambda=0.7,
             Hubble=H, CrossSectionInternal=sig, Seed=55)
hp.set_particles(pos, vel, mass, ids, curtime=np.full(n, A, np.float32))
t0 = time.perf_counter()
hp.predict_collisionless_only(A); hp.force_treebuild(); hp.setup_smoothinglengths_sidm(30)
torch.cuda.synchronize(); t_setup = time.perf_counter() - t0
vmax = hp.getvmax()
t0 = time.perf_counter(); hp.compute_accelerations(1, time=A, vmax=vmax); torch.cuda.synchronize(); t_bh = time.perf_counter() - t0
out = []
t = A
for s in range(3):
    t += 1e-4
    t0 = time.perf_counter(); hp.compute_accelerations(0, time=t, vmax=vmax); hp.advance(time=t); torch.cuda.synchronize()
    c = hp.counters()
    out.append(dict(ms=round((time.perf_counter() - t0) * 1e3, 2), build=round(c.ms_build, 2), walk=round(c.ms_walk, 2), sidm=round(c.ms_sidm, 2),
                    ensure=round(c.ms_ensure, 2), node_int=round(c.node_interactions / n, 1), part_int=round(c.part_interactions / n, 1),
                    repaired=c.ensure_repaired, passes=c.ensure_iterations))
    print(out[-1], flush=True)
acc, ngb = hp.get("Accel", "NgbVelDisp")
assert np.isfinite(acc).all() and ngb.min() >= 28 and ngb.max() <= 32
print(json.dumps(dict(n=n, setup_s=round(t_setup, 2), first_force_bh_s=round(t_bh, 3), steps=out, updates_per_s=round(n / (out[-1]["ms"] * 1e-3)))))
hp.close()
