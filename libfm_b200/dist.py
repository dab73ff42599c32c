"""Multi-GPU plumbing for the row-sharded epoch (host logic, backend-agnostic).

The path shards by ROW: rank g owns rows [g*N/G, (g+1)*N/G) of the training set
and a full replica of w0|w|V; the only exchange is ONE all-reduce of the packed
parameter buffer per epoch followed by a 1/G scale (parameter averaging).  The
reference has no multi-device path; this is the SURVEY.md section 8(e) design.
torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os


def shard_bounds(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, exhaustive, non-overlapping row ranges; sizes differ by <= 1."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return n_rows * rank // world, n_rows * (rank + 1) // world


def shard(data, world: int, rank: int):
    lo, hi = shard_bounds(data.num_cases, world, rank)
    return data.rows(lo, hi)


def env_world() -> tuple[int, int, int]:
    """(world, rank, local_rank) from the torchrun environment."""
    g = lambda k, d: int(os.environ.get(k, d))  # noqa: E731
    return g("WORLD_SIZE", 1), g("RANK", 0), g("LOCAL_RANK", 0)


def allreduce_mean_(params, world: int):
    """In-place parameter averaging of a packed state tensor across ranks."""
    if world <= 1:
        return params
    import torch.distributed as dist
    dist.all_reduce(params, op=dist.ReduceOp.SUM)
    params.mul_(1.0 / world)
    return params


def sum_metrics(sq: float, ab: float, ok: int, n: int, world: int, device=None):
    """Combine per-shard evaluate() sums into global ones (RMSE / accuracy inputs)."""
    if world <= 1:
        return sq, ab, ok, n
    import torch
    import torch.distributed as dist
    t = torch.tensor([sq, ab, float(ok), float(n)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0]), float(t[1]), int(round(float(t[2]))), int(round(float(t[3])))
