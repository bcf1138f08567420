"""Sharding over the GPUs of one box: candidates of a sweep, and -- the same idea one level up --
the independent hyper-parameter restarts of a batched LML evaluation (BASELINE config 5) and the independent tasks of
a batched SVGP scan.  Each has exactly one collective.

Every rank holds a replica of the fitted GP and scores a contiguous shard of the global candidate index
range; the only bytes that cross NVLink are one all-gather of ``topk`` packed (value, index) pairs per rank,
followed by a local lexicographic reduce (value desc, global index asc) that is identical on every rank
and for every GPU count.  The reference has no multi-GPU path; the sharded pool replaces
its single-device chunk loop (optimization/Bayesian7.py:664-672).
"""
from __future__ import annotations

from typing import Tuple

import torch

_I64_MAX = torch.iinfo(torch.int64).max


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [first, first+count) of [0, total): first = rank * ceil(total / world)."""
    per = -(-int(total) // int(world))
    first = min(rank * per, total)
    return first, max(0, min(per, total - first))


def merge_topk(vals: torch.Tensor, idx: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Reduce stacked per-shard lists (any shape, flattened) to the global top-k by (value desc, index asc).

    Empty slots are (-inf, -1) as written by bo_sweep; they sort last.  NaN never appears (the kernel
    maps NaN scores to -inf)."""
    v = vals.reshape(-1).clone()
    i = idx.reshape(-1).clone()
    i[i < 0] = _I64_MAX
    o1 = torch.sort(i, stable=True).indices                    # secondary key: index ascending
    v, i = v[o1], i[o1]
    o2 = torch.sort(v, descending=True, stable=True).indices   # primary key: value descending (stable)
    v, i = v[o2][:k], i[o2][:k]
    i = torch.where(i == _I64_MAX, torch.full_like(i, -1), i)
    if v.numel() < k:
        pad = k - v.numel()
        v = torch.cat([v, torch.full((pad,), float("-inf"), dtype=v.dtype, device=v.device)])
        i = torch.cat([i, torch.full((pad,), -1, dtype=i.dtype, device=i.device)])
    return v, i


def allgather_topk(vals: torch.Tensor, idx: torch.Tensor, k: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The one collective of a sharded sweep: all-gather of k packed (value bits, index) int64 pairs per rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_topk(vals, idx, k)
    world = dist.get_world_size(group)
    packed = torch.stack([vals.contiguous().view(torch.int64), idx.contiguous()], dim=1)      # [k, 2] int64
    out = torch.empty((world * packed.shape[0], 2), dtype=torch.int64, device=packed.device)   # concatenated layout
    dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
    return merge_topk(out[..., 0].contiguous().view(torch.float64), out[..., 1].contiguous(), k)


def sharded_sweep(engine, acq, best_f, beta, sobol, total: int, topk: int, rank: int = 0, world: int = 1,
                  group=None, host: bool = False):
    """Score this rank's shard of a Sobol pool of ``total`` candidates and return the GLOBAL top-k."""
    first, count = shard_range(total, rank, world)
    # the contraction mode is resolved on the GLOBAL pool size and pinned, so every rank takes the same path whatever
    # its shard size (per-candidate values are then bit-identical for every GPU count)
    saved = engine.sweep_mode
    engine.set_sweep_mode(engine.resolve_sweep_mode(total))
    try:
        if host:
            v, i = engine.sweep_host(acq, best_f, beta, sobol=sobol, first_index=first, count=count, topk=topk)
            v, i = v.to(engine.device), i.to(engine.device)
        else:
            v, i = engine.sweep(acq, best_f, beta, sobol=sobol, first_index=first, count=count, topk=topk)
    finally:
        engine.set_sweep_mode(saved)
    if world == 1:
        return v, i
    return allgather_topk(v, i, topk, group)


def sharded_lml_grad(engine, X, y, thetas, kernel="matern52", mean: float = 0.0, rank: int = 0, world: int = 1, group=None):
    """Batched LML + gradient with the R restarts split contiguously over the ranks (they are independent
    factorisations: nothing but the results crosses NVLink).  One all-gather of [ceil(R/world), p + 2] doubles per rank;
    every rank returns the full (lml[R], grad[R,p], status[R]) on the host, like ``GPEngine.lml_grad_batched``."""
    th = torch.as_tensor(thetas, dtype=torch.float64, device="cpu")
    th = th.reshape(-1, th.shape[-1]) if th.ndim > 1 else th.reshape(1, -1)
    R, p = th.shape
    if world == 1:
        return engine.lml_grad_batched(X, y, th, kernel, mean)
    import torch.distributed as dist
    per = -(-R // world)
    first, count = shard_range(R, rank, world)
    buf = torch.zeros(per, p + 2, dtype=torch.float64)
    buf[:, 0] = float("-inf")
    if count > 0:
        lml, grad, status = engine.lml_grad_batched(X, y, th[first:first + count], kernel, mean)
        buf[:count, 0] = torch.as_tensor(lml, dtype=torch.float64).cpu()
        buf[:count, 1] = torch.as_tensor(status, dtype=torch.float64).cpu()
        buf[:count, 2:] = torch.as_tensor(grad, dtype=torch.float64).cpu()
    dev = engine.device if dist.get_backend(group) == "nccl" else torch.device("cpu")
    out = torch.empty(world * per, p + 2, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf.to(dev).contiguous(), group=group)
    out = out.cpu()[:R] if per * world >= R else out.cpu()
    # rank r's rows sit at [r * per, r * per + count_r): with contiguous shards of ceil(R / world) that is already restart order
    return out[:, 0].contiguous(), out[:, 2:].contiguous(), out[:, 1].to(torch.int32)


def task_shard(num_tasks: int, rank: int, world: int):
    """Tasks of a batched SVGP handled by this rank (contiguous split)."""
    first, count = shard_range(num_tasks, rank, world)
    return list(range(first, first + count))


def allreduce_score(score: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the per-rank partial variance scores of a task-sharded SVGP scan (one all-reduce of N doubles)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(score, op=dist.ReduceOp.SUM, group=group)
    return score
