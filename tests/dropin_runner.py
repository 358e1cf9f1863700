"""helper for tests/test_gpu_dropin.py: runs the reference DRIVER (accel.c compute_accelerations,
init.c setup_smoothinglengths_sidm, timeline.c ...) either with its own CPU hot path
(kind=diag) or linked against the product's shim + libsidm_b200.so (kind=b200), on the same
seeded input, and writes the particle fields the path owns."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "sidm-nbody_b200"))


def main(kind, out, n):
    import refdrv
    from sidm_b200 import ic
    pos, vel, mass, ids = ic.hernquist(n, seed=9)
    os.chdir(tempfile.mkdtemp())
    R = refdrv.Reference(kind)
    R.setup(n, CrossSectionInternal=0.0)
    R.init_rand(55)
    R.set_particles(pos, vel, mass, ids)
    R.treebuild()
    R.setup_smoothinglengths_sidm(30)                  # init.c:431 (reference code; its k-NN + count calls hit the path)
    h0, ngb0 = R.get("HSML"), R.get("NGB")
    R.all_active(0.0, 0.0)
    R.getvmax()
    R.compute_accelerations(1)                         # accel.c:27, start-up forces (BH, OldAcc = 0)
    acc1, old1 = R.get("ACCEL"), R.get("OLDACC")
    rng = np.random.default_rng(4)
    R.set("HSML", (h0 * rng.choice(np.array([1, 1, 1, 0.8, 1.3], np.float32), n)).astype(np.float32))
    R.all_active(0.0, 0.01)
    R.compute_accelerations(0)                         # gravity (relative criterion) + sidm + ensure_neighbours
    res = dict(h0=h0, ngb0=ngb0, acc1=acc1, old1=old1, acc2=R.get("ACCEL"), old2=R.get("OLDACC"),
               h2=R.get("HSML"), ngb2=R.get("NGB"), dvel=R.get("DVEL"), pospred=R.get("POSPRED"),
               nactive=len(R.active()))
    if kind != "b200":                                 # potential.c:18 (the symbol-by-symbol shim leaves potential.c's CPU walk out)
        R.compute_potential()
        res["pot"] = R.get("POT")
        res["sys"] = R.global_quantities()             # global.c:18 (b200f: the shim's, from the device state)
        snap = R.savepositions(5, os.getcwd(), mass_table=[0, float(mass[0]), 0, 0, 0, 0], hubble_param=0.7)   # io.c:16
        res["snap"] = np.frombuffer(open(snap, "rb").read(), np.uint8)
    np.savez(out, **res)


def main_run(kind, out, n, k=60):
    """the reference's own main loop (run.c:34-150) with individual time steps: start-up as init.c:120-185, then k
    iterations of find_next_time / compute_accelerations(0) / advance / find_timesteps(0) on small active sets"""
    import refdrv
    from sidm_b200 import ic
    pos, vel, mass, ids = ic.hernquist(n, seed=10)
    os.chdir(tempfile.mkdtemp())
    R = refdrv.Reference(kind)
    R.setup(n, CrossSectionInternal=0.0)      # no scatterings: the two runs can be compared particle by particle
    R.init_rand(55)
    R.set_particles(pos, vel, mass, ids)
    R.treebuild()
    R.setup_smoothinglengths_sidm(30)
    R.all_active(0.0, 0.0)
    R.getvmax()
    R.compute_accelerations(1)
    ts = dict(crit=0, eta=0.02, velscale=10.0, probtol=0.2, dyntol=0.05, dtmax=0.02, dtmin=0.0)
    R.find_timesteps(2, **ts)                          # init.c:177: first steps + construct_timetree()
    mp0 = R.get("MAXPRED")
    t, na = R.run_steps(k)
    np.savez(out, maxpred0=mp0, time=t, nactive=na, pos=R.get("POS"), vel=R.get("VEL"), curtime=R.get("CURTIME"),
             maxpred=R.get("MAXPRED"), hsml=R.get("HSML"), ngb=R.get("NGB"))


def main_sct(kind, out, n):
    """-DSCATTERLOG through the drop-in (sidm.c:96-104, 571-601): one all-active compute_accelerations(0) with a large cross
    section; the records the run appends to sct_<snapshot count>.<task> in its working directory"""
    import refdrv
    from sidm_b200 import ic
    pos, vel, mass, ids = ic.hernquist(n, seed=11)
    os.chdir(tempfile.mkdtemp())
    R = refdrv.Reference(kind)
    R.setup(n, CrossSectionInternal=400.0)
    R.init_rand(55)
    R.set_particles(pos, vel, mass, ids)
    R.treebuild()
    R.setup_smoothinglengths_sidm(30)
    R.all_active(0.0, 0.0)
    R.getvmax()
    R.compute_accelerations(1)
    R.all_active(0.0, 0.01)
    R.compute_accelerations(0)
    log = refdrv.read_scatlog("sct_000.0")
    np.savez(out, log=log, dvel=R.get("DVEL"), ids=R.get("ID"), hsml=R.get("HSML"))


if __name__ == "__main__":
    if len(sys.argv) > 4 and sys.argv[4] == "sct":
        main_sct(sys.argv[1], sys.argv[2], int(sys.argv[3]))
    elif len(sys.argv) > 4 and sys.argv[4] == "run":
        main_run(sys.argv[1], sys.argv[2], int(sys.argv[3]))
    else:
        main(sys.argv[1], sys.argv[2], int(sys.argv[3]))
