"""Multi-GPU plumbing of the hot path: reads are sharded by record, every rank holds a full
index replica, and the only collective is ONE all-reduce (SUM) of the Readstats counter vector
(include/readstats.hpp:77-84) after the last batch -- torch.distributed over NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_reads: int, rank: int, world: int) -> tuple:
    """Record-aligned contiguous shard [lo, hi) of rank -- the analogue of the reference's Readfeed cutting a
    reads file into `num_splits` record-aligned ranges (src/sortmerna/readfeed.cpp:1253-1277): sizes differ by at most 1."""
    base, rem = divmod(n_reads, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def counter_vector(counters: dict, matched, names) -> np.ndarray:
    return np.array([int(counters[k]) for k in names] + [int(x) for x in matched], dtype=np.int64)


def allreduce_counters(vec: np.ndarray, device=None) -> np.ndarray:
    """The path's single collective.  No-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return vec
    t = torch.from_numpy(np.ascontiguousarray(vec))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allreduce_max(vec: np.ndarray, device=None) -> np.ndarray:
    """max over ranks (device-side timings)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return vec
    t = torch.from_numpy(np.ascontiguousarray(vec))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()
