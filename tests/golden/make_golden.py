"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile) on small seeded inputs.  Run in the build container:

    python tests/golden/make_golden.py            # hernquist3k.npz: one particle type, tree / forces / neighbours / sidm
    python tests/golden/make_golden.py global     # global3k.npz + types3k.npz: three types - SysState, snapshot file hash,
                                                  # forces, potentials, start-up smoothing lengths

The reference ships no golden vectors of its own; these fixtures pin the
oracle restatement (oracle/*.c) and, through it, the CUDA path, on machines where
/root/reference and oracle/_ref are absent."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "sidm-nbody_b200"))

import refdrv  # noqa: E402
from sidm_b200 import ic  # noqa: E402

N = 3000
SIGMA = 208.9      # 100 cm^2/g: many events in a tiny halo
DT = 0.02


def main():
    pos, vel, mass, ids = ic.hernquist(N, seed=11)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    R = refdrv.Reference("diag")
    R.setup(N, CrossSectionInternal=SIGMA)
    R.init_rand(55)
    R.set_particles(pos, vel, mass, ids)
    R.treebuild()
    nd = R.dump_nodes()
    chain = []
    nxt = R.dump_next()
    p = int(nd["partind"][0])
    while p >= 0:
        chain.append(p)
        p = int(nxt[p])
    idx = np.arange(0, N, 7, dtype=np.int32)
    acc_bh, cost_bh = R.force_tree(idx)
    direct = R.force_direct(idx)
    full = np.arange(N, dtype=np.int32)
    acc_full, _ = R.force_tree(full, want_cost=False)
    a = acc_full.astype(np.float32)
    oldacc = np.sqrt((a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1] + a[:, 2] * a[:, 2]).astype(np.float64)).astype(np.float32)
    R.set("OLDACC", oldacc)
    acc_rel, cost_rel = R.force_tree(idx)
    # gravity_tree() end to end (first call BH, second relative)
    R.set("OLDACC", np.zeros(N, np.float32))
    R.all_active(0.0, 0.0)
    R.gravity_tree()
    g1_acc, g1_old = R.get("ACCEL"), R.get("OLDACC")
    R.gravity_tree()
    g2_acc, g2_old = R.get("ACCEL"), R.get("OLDACC")
    # smoothing lengths, k-NN, neighbour lists
    R.setup_smoothinglengths_sidm(30)
    hsml, ngb0 = R.get("HSML"), R.get("NGB")
    knn = np.array([R.ngb_treefind(pos[i], 30) for i in idx], np.float32)
    lists = [R.ngb_variable(pos[i], hsml[i])[0] for i in idx]
    lmax = max(len(x) for x in lists)
    nlist = np.full((len(idx), lmax), -1, np.int32)
    for k, x in enumerate(lists):
        nlist[k, :len(x)] = x
    # one sidm() call with the RNG stream logged
    R.all_active(0.0, DT / 2)
    t_sidm = R.time
    vmax = R.getvmax()
    R.rng_log_begin(20 * N)
    R.sidm()
    draws = R.rng_log_end()
    dvel1, ngb1 = R.get("DVEL"), R.get("NGB")
    log = refdrv.read_scatlog("sct_000.0")
    # then the repair loop (h perturbed so that it has work to do), sigma -> 0 so no RNG dependence
    os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "hernquist3k.npz"), pos=pos, vel=vel, mass=mass, ids=ids,
                        node_center=nd["center"], node_len=nd["len"], node_mass=nd["mass"], node_s=nd["s"],
                        node_Q=np.concatenate([nd["Q"], nd["P"][:, None]], axis=1), node_oc=nd["oc"],
                        node_bmax2=nd["bmax2"], node_count=nd["count"], chain=np.array(chain, np.int32),
                        idx=idx, acc_bh=acc_bh, cost_bh=cost_bh, direct=direct, oldacc=oldacc, acc_rel=acc_rel,
                        cost_rel=cost_rel, g1_acc=g1_acc, g1_old=g1_old, g2_acc=g2_acc, g2_old=g2_old,
                        hsml=hsml, ngb0=ngb0, knn=knn, nlist=nlist, t_sidm=t_sidm, vmax=vmax, draws=draws,
                        dvel1=dvel1, ngb1=ngb1, log_id1=log["id1"], log_id2=log["id2"], log_dv=log["dv"],
                        sigma=SIGMA, dt=DT)
    print("wrote hernquist3k.npz:", len(nd["len"]), "nodes,", len(log), "scatter events,", len(draws), "draws")


def main_global():
    """global3k.npz: compute_potential() + compute_global_quantities_of_system() (potential.c:18, global.c:18) on the
    same halo half a step into a run (PosPred != Pos), with three particle types of different softening."""
    pos, vel, mass, ids = ic.hernquist(N, seed=11)
    types = np.random.default_rng(12).choice(np.array([1, 2, 3], np.int32), N, p=[0.6, 0.3, 0.1]).astype(np.int32)
    eps = {1: 0.3, 2: 0.6, 3: 0.2}
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    R = refdrv.Reference("diag")
    R.setup(N, CrossSectionInternal=0.0)
    for t, e in eps.items():
        R.set_softening(t, e)
    R.set_particles(pos, vel, mass, ids)
    R.set("TYPE", types)
    R.all_active(0.0, 0.0)
    R.getvmax()
    R.compute_accelerations(1)
    R.compute_potential()
    sys_state = R.global_quantities()
    # the snapshot file of this state (savepositions(), io.c:16): type 2 has tabulated masses
    import hashlib
    mt = [0, 0, float(2 * mass[0]), 0, 0, 0]
    snap = open(R.savepositions(3, os.getcwd(), mass_table=mt, hubble_param=0.7), "rb").read()
    # three trees: forces (both criteria), interaction counts, raw potentials, start-up smoothing lengths
    oldacc = R.get("OLDACC")
    posp, velp, massp, potp = R.get("POSPRED"), R.get("VELPRED"), R.get("MASS"), R.get("POT")
    idx = np.arange(0, N, 7, dtype=np.int32)
    R.treebuild()
    R.set("OLDACC", np.zeros(N, np.float32))
    acc_bh, cost_bh = R.force_tree(idx)
    R.set("OLDACC", oldacc)
    acc_rel, cost_rel = R.force_tree(idx)
    pot_raw = R.potential(idx)
    R.setup_smoothinglengths_sidm(30)
    hsml, ngb = R.get("HSML"), R.get("NGB")
    os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "types3k.npz"), idx=idx, acc_bh=acc_bh, cost_bh=cost_bh, acc_rel=acc_rel, cost_rel=cost_rel,
                        pot_raw=pot_raw, hsml=hsml, ngb=ngb)
    np.savez_compressed(os.path.join(HERE, "global3k.npz"), types=types, eps=np.array([0, 0.3, 0.6, 0.2, 0, 0]),
                        pospred=posp, velpred=velp, mass=massp, oldacc=oldacc,
                        pot=potp, sys=sys_state,
                        ids=R.get("ID"), snap_mass_table=np.array(mt), snap_time=R.time, snap_len=len(snap),
                        snap_sha256=hashlib.sha256(snap).hexdigest(), snap_head=np.frombuffer(snap[:264], np.uint8),
                        snap_omega0=R.cfg["Omega0"])
    print("wrote global3k.npz: E_kin", sys_state[1], "E_pot", sys_state[2])


if __name__ == "__main__":
    if "global" in sys.argv[1:]:
        main_global()
    else:
        main()
