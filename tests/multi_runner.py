"""helper for tests/test_gpu_multi.py: runs two steps of the hot path as rank RANK of WORLD
(gloo = ranks share cuda:0 and stage the all-gather through the host; nccl = one GPU each)
and writes the particle fields the path owns."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sidm-nbody_b200"))


def main(out, n, backend):
    import torch
    import torch.distributed as dist
    from sidm_b200 import HotPath, ic
    from sidm_b200.multi import Sharder
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = int(os.environ.get("LOCAL_RANK", "0")) if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group(backend)
    pos, vel, mass, ids = ic.hernquist(n, seed=13)
    multi_type = os.environ.get("B200_TYPES", "0") == "1"          # three particle types: one tree per type
    types = np.random.default_rng(17).choice(np.array([1, 2, 3], np.int32), n, p=[0.5, 0.3, 0.2]).astype(np.int32) if multi_type else np.ones(n, np.int32)
    kw = dict(SofteningTable=[0, 0.3, 0.6, 0.2, 0, 0]) if multi_type else {}
    hp = HotPath(n, device=dev, CrossSectionInternal=400.0, Seed=7, **kw)
    sh = Sharder(hp, world, rank)
    hp.set_option("shard_min_work", int(os.environ.get("B200_SHARD_MIN_WORK", "1")))   # shard even this small problem
    hp.set_option("compact_exchange", int(os.environ.get("B200_COMPACT", "1")))
    hp.set_particles(pos, vel, mass, ids)
    if multi_type:
        hp.set_field("ptype", types)
    hp.predict_collisionless_only(0.0)
    hp.force_treebuild()
    hp.setup_smoothinglengths_sidm(30)
    vmax = hp.getvmax()
    sh.compute_accelerations(1, time=0.0, vmax=vmax)
    acc1, old1 = hp.get("Accel", "OldAcc")
    dt = 0.004
    t = 0.0
    res = {}
    for step in range(2):
        sh.compute_accelerations(0, time=t + dt / 2, vmax=vmax)
        c = hp.counters()
        res[f"acc{step}"], res[f"old{step}"], res[f"dvel{step}"], res[f"h{step}"], res[f"ngb{step}"] = hp.get(
            "Accel", "OldAcc", "dVel", "HsmlVelDisp", "NgbVelDisp")
        res[f"sct{step}"] = np.array([c.sct_ntot, c.sct_pass1, c.sct_scattered, c.sct_rejected, c.ensure_iterations])
        hp.advance(time=t + dt / 2)
        t += dt
    # the host array-of-structs path: own rows over PCIe, replication by all-gather, own rows back
    from sidm_b200 import capi
    velh = hp.peek("velh", np.float32, (n, 4))
    aos = np.zeros(n, capi.PARTICLE_DTYPE)
    aos["Pos"] = hp.peek("pos0", np.float32, (n, 3)); aos["PosPred"] = aos["Pos"]
    aos["Vel"] = velh[:, :3]; aos["VelPred"] = aos["Vel"]; aos["HsmlVelDisp"] = velh[:, 3]
    aos["Mass"] = mass; aos["ID"] = ids; aos["Type"] = types; aos["Potential"] = 7.0; aos["ForceFlag"] = 3
    aos["CurrentTime"] = hp.peek("curtime", np.float32, (n,))
    aos["Accel"], aos["OldAcc"], aos["NgbVelDisp"] = hp.get("Accel", "OldAcc", "NgbVelDisp")
    back = aos.copy()
    hp.bind_particles(aos, pin=False)
    sh.upload()
    sh.compute_accelerations(0, time=t + dt / 2, vmax=vmax)
    sh.download(into=back)
    first, cnt, per = sh.rows() if world > 1 else (0, n, n)
    lo, hi = 0, -(-n // 8)                             # rows that rank 0 owns at every world size up to 8
    for f in ("Accel", "OldAcc", "dVel", "HsmlVelDisp", "NgbVelDisp", "PosPred", "Potential", "ForceFlag"):
        res["aos_" + f] = back[f][lo:hi].copy()
    if world > 1 and rank == 0:
        assert np.array_equal(back["Accel"][first + cnt:], aos["Accel"][first + cnt:]), "rows of other ranks must stay untouched"
    # statistics on every rank (replicated state), snapshot by rank 0 only
    res["sys"] = sh.compute_global_quantities_of_system().flat()
    snap = out + ".snap"
    sh.savepositions(snap, time=t, mass_table=[0, 0, 0.25, 0, 0, 0])
    assert os.path.exists(snap)
    if rank == 0:
        res["snap"] = np.frombuffer(open(snap, "rb").read(), np.uint8)
        np.savez(out, acc_start=acc1, old_start=old1, **res)
    hp.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3])
