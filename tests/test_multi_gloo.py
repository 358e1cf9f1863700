"""CPU, world_size 2 over gloo: the host-side sharding logic of sidm_b200/multi.py - which list
positions a rank owns, the layout of the all-gather buffer and its unpacking - exercised with a
real collective.  (The CUDA kernels k_shard_select / k_*_pack / k_*_unpack use the same index
arithmetic; the 2-GPU run itself is checked by tests/test_gpu_multi.py.)"""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nt, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "sidm-nbody_b200"))
    import torch
    import torch.distributed as dist
    from sidm_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    work = rng.permutation(nt).astype(np.int64)            # a sorted work list (slot ids)
    truth = (work * 3 + 1).astype(np.float32)              # what the "kernel" computes per entry
    mine = multi.shard_positions(nt, world, rank)
    per_rank = multi.shard_max_blocks(nt, world) * multi.SHARD_BLOCK
    send = torch.zeros(per_rank, dtype=torch.float32)
    send[: len(mine)] = torch.from_numpy(truth[mine])      # own results, packed in block order
    recv = torch.empty(world * per_rank, dtype=torch.float32)
    dist.all_gather_into_tensor(recv, send)
    pos, pr = multi.unpack_positions(nt, world)
    assert pr == per_rank
    out = np.full(nt, np.nan, np.float32)
    ok = pos >= 0
    out[pos[ok]] = recv.numpy()[ok]
    q.put((rank, bool(np.array_equal(out, truth)), len(mine)))
    dist.destroy_process_group()


@pytest.mark.parametrize("nt", [1, 31, 32, 33, 63, 64, 65, 1000, 4097])
def test_shard_roundtrip_world2(nt):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, nt, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(r[1] for r in res)
    assert sum(r[2] for r in res) == nt


def test_shard_partition_properties():
    import sys
    from sidm_b200 import multi
    for nt in (0, 1, 64, 65, 12345):
        for world in (1, 2, 4, 8):
            allpos = np.concatenate([multi.shard_positions(nt, world, r) for r in range(world)])
            assert sorted(allpos.tolist()) == list(range(nt))          # a partition of the list
            sizes = [len(multi.shard_positions(nt, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= multi.SHARD_BLOCK                         # balanced to one block
            assert multi.buffer_bytes(nt, world) >= max(sizes) * multi.SLOT_REC_BYTES
