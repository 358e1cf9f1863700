"""Sharding of the hot path across the GPUs of one box (one process per GPU).

Replaces the role of the reference's domain.c + hypercube exchanges: the
particle state is replicated on every GPU (1e7 particles = a few GB of 180 GB), every rank
builds the same octree, and the TARGETS are split: rank r walks / scatters the 32-entry
blocks b of every work list sorted along the octant-key order with b % world == r
(interleaved blocks balance the dense centre across ranks and keep each warp's targets
spatial neighbours).  Per-target results are exchanged with one NCCL all-gather per phase
over NVLink; the library packs/unpacks and calls back into `torch.distributed` for the
collective (include/sidm_b200.h: b200_set_shard).  Counter-based per-particle random numbers
make the N-GPU result identical to the 1-GPU result.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_longlong, C.c_void_p)
SLOT_REC_BYTES = 32        # struct SlotRec in csrc/sidm.cu; gravity sends 16-byte float4 records


SHARD_BLOCK = 32      # kShardBlock of csrc/ctx.cuh: the 32 targets one warp of the walk handles


def shard_max_blocks(nt, world):
    return ((nt + SHARD_BLOCK - 1) // SHARD_BLOCK + world - 1) // world


def shard_positions(nt, world, rank):
    """positions j of a sorted work list of length nt that rank owns (mirror of k_shard_select)"""
    k = np.arange(shard_max_blocks(nt, world) * SHARD_BLOCK, dtype=np.int64)
    j = ((k // SHARD_BLOCK) * world + rank) * SHARD_BLOCK + (k % SHARD_BLOCK)
    return j[j < nt]


def unpack_positions(nt, world):
    """for the concatenated all-gather buffer [world][per_rank]: list position of every entry, -1 = padding
    (mirror of k_grav_unpack / k_slot_unpack)"""
    per_rank = shard_max_blocks(nt, world) * SHARD_BLOCK
    g = np.arange(world * per_rank, dtype=np.int64)
    q, k = g // per_rank, g % per_rank
    j = ((k // SHARD_BLOCK) * world + q) * SHARD_BLOCK + (k % SHARD_BLOCK)
    return np.where(j < nt, j, -1), per_rank


def buffer_bytes(n, world, stride=0):
    """size of the send buffer: per-slot records of one rank's share, or - when the host array is
    sharded too (stride = bytes per particle row) - one rank's rows of the array-of-structs"""
    rows = -(-n // world)
    return max(shard_max_blocks(n, world) * SHARD_BLOCK * SLOT_REC_BYTES, rows * stride)


class Sharder:
    def __init__(self, hp, world=1, rank=0, aos_stride=124):
        self.hp, self.world, self.rank = hp, int(world), int(rank)
        self._cb = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            cap = buffer_bytes(hp.params.MaxPart, self.world, aos_stride)
            self.send = torch.empty(cap, dtype=torch.uint8, device="cuda")
            self.recv = torch.empty(cap * self.world, dtype=torch.uint8, device="cuda")
            self.exchanges = 0
            self.bytes = 0

            host_staged = dist.get_backend() == "gloo"      # test harness: several ranks sharing one GPU

            def allgather(nbytes, user):
                try:
                    # order the collective on the stream the library is working on (its main stream or
                    # the SIDM stream): b200_current_stream(), include/sidm_b200.h
                    sp = hp.lib.b200_current_stream()
                    stream = torch.cuda.ExternalStream(sp) if sp else torch.cuda.default_stream()
                    with torch.cuda.stream(stream):
                        if host_staged:
                            out = torch.empty(self.world * nbytes, dtype=torch.uint8)
                            dist.all_gather_into_tensor(out, self.send[:nbytes].cpu())
                            self.recv[: self.world * nbytes].copy_(out)
                        else:
                            dist.all_gather_into_tensor(self.recv[: self.world * nbytes], self.send[:nbytes])
                    self.exchanges += 1
                    self.bytes += int(nbytes)
                    return 0
                except Exception as e:  # pragma: no cover
                    print("all-gather failed:", e, flush=True)
                    return 1

            self._cb = ALLGATHER_FN(allgather)
            rc = hp.lib.b200_set_shard(self.rank, self.world, C.c_void_p(self.send.data_ptr()), C.c_void_p(self.recv.data_ptr()),
                                       C.c_longlong(cap), self._cb, None)
            if rc != 0:
                raise RuntimeError(f"b200_set_shard -> {rc}")
            hp.set_option("shard_overlap", 1)          # the callback above honours b200_current_stream()

    def describe(self):
        if self.world == 1:
            return "1 GPU"
        return (f"{self.world} GPUs: replicated particles + tree, targets split by interleaved 32-particle blocks of the tree order, "
                "NCCL all-gather of accelerations and scatter proposals, SIDM chain on its own stream next to the walk")

    def compute_accelerations(self, mode, time, vmax, active=None):
        self.hp.compute_accelerations(mode, active, time, vmax)

    # statistics and snapshots (global.c:18, io.c:16): the particle state is replicated, so every rank computes the same
    # SysState (deterministic block sums) and ONE rank writes the file - the reference reduces / sends to task 0 instead
    # (global.c:59-66, io.c:90-103,390-470)
    def compute_global_quantities_of_system(self):
        return self.hp.compute_global_quantities_of_system()

    def savepositions(self, path, **kw):
        npart = self.hp.savepositions(path, **kw) if self.rank == 0 else None
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        return npart

    # host array-of-structs sharded over the ranks (the reference's per-rank P[]): rank r owns rows
    # [r*rows, (r+1)*rows); PCIe carries only the own rows, NVLink replicates them
    def rows(self):
        per = -(-self.hp.n // self.world)
        first = min(self.rank * per, self.hp.n)
        return first, min(per, self.hp.n - first), per

    def upload(self):
        if self.world == 1:
            self.hp.upload()
            return
        first, cnt, per = self.rows()
        rc = self.hp.lib.b200_upload_shard(first, cnt, per)
        if rc != 0:
            raise RuntimeError(f"b200_upload_shard -> {rc}")

    def download(self, into=None):
        if self.world == 1:
            self.hp.download(into=into)
            return
        first, cnt, per = self.rows()
        dst = None if into is None else into.ctypes.data_as(C.c_void_p)
        rc = self.hp.lib.b200_download_shard(dst, first, cnt)
        if rc != 0:
            raise RuntimeError(f"b200_download_shard -> {rc}")
