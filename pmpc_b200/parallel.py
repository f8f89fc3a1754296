"""Multi-GPU plumbing: one process per GPU (``torch.distributed``), problems sharded, no data-path
collective (SURVEY.md 8e case 1: independent problems shard naturally).

The reference's counterpart is the ZeroMQ task farm ``pmpc/remote.py:391-452`` (one problem per CPU
worker process over TCP).  Here rank r owns the contiguous slice ``shard_bounds(B, r, world)`` of the
problem batch and runs it on its own B200; only the per-rank wall/device time is reduced (max) for
reporting.  Particle sharding of ONE consensus problem (8e case 2) needs a per-ADMM-iteration
all-reduce of the Nc*m consensus terms and is not in this round (DESIGN.md "next").
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Balanced contiguous partition: the first ``total % world`` ranks get one extra item."""
    assert 0 <= rank < world and total >= 0
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_process_group(backend: str = "nccl"):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    rank, world, local = env_rank_world()
    if world == 1:
        return None
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return dist


def reduce_max(value: float, dist=None, device=None) -> float:
    """max over ranks of a scalar (timing): the job finishes when the slowest rank does."""
    if dist is None:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float, dist=None, device=None) -> float:
    if dist is None:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_objects(obj, dist=None) -> List:
    """All ranks' python objects on every rank, in rank order (results of sharded batches)."""
    if dist is None:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def solve_problems_sharded(problems: Sequence[dict], dist=None, device: int = -1):
    """Rank-local slice of ``problems`` through the batched device path; every rank returns the full,
    ordered result list (gathered)."""
    from .batch import solve_batched

    rank, world = (0, 1) if dist is None else (dist.get_rank(), dist.get_world_size())
    lo, hi = shard_bounds(len(problems), rank, world)
    local = solve_batched(list(problems[lo:hi]), device=device)
    parts = gather_objects(local, dist)
    return [r for part in parts for r in part]
