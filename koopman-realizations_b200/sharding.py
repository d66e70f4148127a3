"""Snapshot sharding for the multi-GPU fit.

The fit shards naturally over snapshots: G = sum_r Px_r' Px_r, C = sum_r Px_r' Py_r.  Rank r owns a
contiguous slice of the snapshot pairs, accumulates its partial Gram tiles on its own GPU, and ONE
all-reduce (sum) of the packed accumulator makes every rank hold the full (G, C); the solve is then
replicated (deterministic, no broadcast).  There is no other data-path collective.

The collective goes through `torch.distributed` (NCCL on GPUs; gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(M, rank, world):
    """Contiguous, balanced slice [lo, hi) of M snapshot pairs owned by `rank` (first M % world ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(int(M), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_sum_(tensor, group=None):
    """In-place sum over ranks of the packed partial-Gram accumulator; no-op for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def budgets_for_rank(nt, rank, world):
    """Round-robin split of a lasso vector over ranks (indices of the budgets rank solves)."""
    return np.arange(rank, nt, world)


class DeviceArrayView:
    """Zero-copy view of a raw device pointer for torch.as_tensor (via __cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
