"""Sharding of a batch of independent problem instances over ranks.

The Newton/KKT path has no data-path exchange: instances are partitioned contiguously, each rank (one process per GPU)
solves its own, and the only collective is the sum of the four convergence counters (running, converged, gave up,
error).  On the GPU the library does that all-reduce itself over NCCL, in-stream (cb200_allreduce_counts); the
torch.distributed variant below is the host-side equivalent (any backend: gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition: the first `total % world` ranks own one instance more.  Returns [begin, end)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("invalid rank/world")
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_counts(counts, group=None) -> dict:
    """Sum the per-rank convergence counters over the process group (no-op without torch.distributed)."""
    keys = ("running", "converged", "gave_up", "error")
    v = np.array([int(counts[k]) for k in keys], dtype=np.int64)
    try:
        import torch
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            t = torch.from_numpy(v.copy())
            if dist.get_backend(group) == "nccl":
                t = t.cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            v = t.cpu().numpy()
    except ImportError:
        pass
    return {k: int(x) for k, x in zip(keys, v)}


def solve_sharded(problems, rank: int, world: int, *, binding=None, device: int = 0, max_steps: int = 400,
                  check_every: int = 4, group=None):
    """Solve this rank's shard of `problems` (LQ-conic instances sharing one pattern) on the device and return
    (local BatchKKT, global counters, [begin, end)).  The loop stops when the all-reduced `running` count is zero."""
    from .solver import BatchKKT
    begin, end = shard_range(len(problems), rank, world)
    mine = problems[begin:end]
    k = None
    local = dict(running=0, converged=0, gave_up=0, error=0)
    if mine:
        k = BatchKKT(mine[0], batch=len(mine), device=device, binding=binding)
        k.load_lq(mine)
        k.initialize(np.stack([P.x0 for P in mine]))
        k.lq_begin()
        local = dict(running=len(mine), converged=0, gave_up=0, error=0)
    done = 0
    glob = allreduce_counts(local, group)
    while glob["running"] > 0 and done < max_steps:
        if k is not None and local["running"] > 0:
            k.lq_step(check_every)
            c = k.allreduce_counts()           # local counts (no NCCL communicator attached to this handle)
            local = {kk: c[kk] for kk in local}
        done += check_every
        glob = allreduce_counts(local, group)
    return k, glob, (begin, end)
