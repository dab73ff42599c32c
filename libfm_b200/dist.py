"""Multi-GPU plumbing for the row-sharded epoch (host logic, backend-agnostic).

The path shards by ROW: rank g owns rows [g*N/G, (g+1)*N/G) of the training set
and a full replica of w0|w|V; the only exchange is ONE combine of the packed
parameter buffer per epoch: the plain mean, or the mean-field weighted delta sum
(combine_meanfield_) that bench.py's NCCL path and the NVLink peer kernel use.  The
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


def meanfield_gamma(u, world: int):
    """gamma = (1 - (1-s)^G) / (G s) with 1 - s = exp(-u): the factor that makes G summed shard-steps of
    relative size s equal G such steps taken one after the other on a quadratic (-> 1 as u -> 0: sum;
    -> 1/G as u -> inf: average).  Same closed form as fm_peer.cu::mf_gamma."""
    import torch
    u = torch.as_tensor(u)
    small = u <= 1e-6
    us = torch.where(small, torch.ones_like(u), u)
    g = -torch.expm1(-world * us) / (world * -torch.expm1(-us))
    return torch.where(small, torch.ones_like(u), g)


def combine_meanfield_(params, theta0, counts, n_rows: int, layout: dict, lr: float, reg0: float = 0.0,
                       regw: float = 0.0, regv: float = 0.0, world: int = 1):
    """The per-epoch exchange of the row-sharded path over torch.distributed (NCCL on GPUs, gloo on CPUs):
    theta = theta0 + gamma_i * sum_g (theta_g - theta0), in place on `params` (this rank's packed fp32 state
    [w0,0,0,0 | w (stride ws) | V[n][kp]] after its shard-epoch).  `theta0` = the common state the epoch started
    from, `counts` = this shard's per-feature occurrence counts [n], `layout` = dict(off_w, ws, off_v, kp, n).
    Restates fm_peer_meanfield_kernel (libfm_b200/csrc/fm_peer.cu) -- the NVLink peer-memory kernel is the
    fast path, this is the fallback where peer mapping is unavailable."""
    if world <= 1:
        return params
    import torch
    import torch.distributed as dist
    off_w, ws, off_v, kp, n = (int(layout[k]) for k in ("off_w", "ws", "off_v", "kp", "n"))
    delta = params - theta0
    dist.all_reduce(delta, op=dist.ReduceOp.SUM)
    cnt = counts.to(torch.float32).clone()
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    cnt /= world
    rows = torch.tensor([float(n_rows)], dtype=torch.float32, device=params.device)
    dist.all_reduce(rows, op=dist.ReduceOp.SUM)
    rows /= world
    v0 = theta0[off_v:off_v + n * kp]
    hv = float((v0.double() ** 2).sum()) / max(n, 1)  # mean squared factor-row norm of theta0
    gamma = torch.zeros_like(params)
    gamma[0] = meanfield_gamma(lr * (1.0 + reg0) * rows, world)[0]
    gamma[off_w:off_w + n * ws:ws] = meanfield_gamma(lr * (1.0 + regw) * cnt, world)
    gamma[off_v:off_v + n * kp] = meanfield_gamma(lr * (hv + regv) * cnt, world).repeat_interleave(kp)
    params.copy_(theta0 + gamma * delta)
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
