"""Captures the active-particle lists of the UNMODIFIED reference's own time line (timeline.c / timestep.c, run.c:34-150) for
BASELINE config C5 (gravothermal-collapse halo, sigma/m = 10 cm^2/g, individual time steps) at reduced N, as SURVEY.md 8d
describes: bench.py --config C5 replays these lists, scaled to N = 4e6 by radius rank, so that the GPU path sees the active-set
sizes and the spatial distribution of a real run.  Run in this container (needs oracle/_ref built from /root/reference):

    python tests/golden/make_c5_timeline.py        ->  tests/golden/c5_timeline.npz

Stored per iteration: All.Time and the radius ranks (0 = innermost) of the active particles."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "sidm-nbody_b200"))

NRED, ITER = 32768, 96
C5_IC = dict(seed=5, rho0=1.49e-4 * 27, rs=11.14356 / 3)          # the sample halo, three times as concentrated (same mass)


def main():
    import refdrv
    from sidm_b200 import ic
    pos, vel, mass, ids = ic.nfw(NRED, **C5_IC)
    sigma = ic.cross_section_internal(10.0)
    os.chdir(tempfile.mkdtemp())
    devnull = os.open(os.devnull, os.O_WRONLY)
    keep = os.dup(1)
    os.dup2(devnull, 1)
    R = refdrv.Reference("fast")
    R.setup(NRED, CrossSectionInternal=sigma, TreeUpdateFrequency=0.1)
    R.init_rand(55)
    R.set_particles(pos, vel, mass, ids)
    R.treebuild()
    R.setup_smoothinglengths_sidm(30)
    R.all_active(0.0, 0.0)
    R.getvmax()
    R.compute_accelerations(1)
    # parameter.txt:114-120: velocity-scale criterion, SIDM and G*rho limits, MaxSizeTimestep 0.1
    ts = dict(crit=1, eta=0.005, velscale=0.66, probtol=0.2, dyntol=0.004, dtmax=0.1, dtmin=0.0)
    R.find_timesteps(2, **ts)
    rank = np.empty(NRED, np.int32)
    rank[np.argsort((pos.astype(np.float64) ** 2).sum(1), kind="stable")] = np.arange(NRED, dtype=np.int32)
    times, offs, ranks = [], [0], []
    for s in range(ITER):
        t, na = R.run_steps(1)
        act = R.active()
        assert len(act) == na[0]
        times.append(t[0]); ranks.append(np.sort(rank[act])); offs.append(offs[-1] + len(act))
    os.dup2(keep, 1)
    out = os.path.join(HERE, "c5_timeline.npz")
    np.savez_compressed(out, nred=NRED, time=np.array(times), offsets=np.array(offs, np.int64), ranks=np.concatenate(ranks).astype(np.int32),
                        ic=np.array([C5_IC["seed"], C5_IC["rho0"], C5_IC["rs"]]))
    na = np.diff(offs)
    print(f"{out}: {ITER} iterations, active per iteration min {na.min()} median {int(np.median(na))} max {na.max()} of {NRED}; dt median {np.median(np.diff(times)):.3e}")


if __name__ == "__main__":
    main()
