"""Trajectory sharding across the GPUs of one box (one process per GPU, torch.distributed/NCCL).

Trajectories are independent IVPs; the only couplings are
  1. the shared weights  -> dL/dtheta is a sum over trajectories   -> all-reduce(sum) of the ODE-func grads
  2. posterior() mean/std over all (beta, gamma) rows               -> all-reduce(sum) of the side sums
  3. ||tracker||_F                                                   -> same 8-double message as 2.
  4. the NLL's mean/std over the SAMPLE axis (lib/train_functions.py:82-83), outside the hot path:
     kept local by sharding along the `batch` sub-index of the flat axis B = n_samples * batch
     (flat index s * batch + b, lib/models.py:23), so every rank holds ALL samples of its b's.

There is no per-step communication: each rank runs the fused kernels on its shard; messages are
one 8-double all-reduce per forward and one n_theta-float all-reduce per backward.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def batch_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of the `batch` sub-index: rank r owns b in [lo, hi)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_flat(x: torch.Tensor, n_samples: int, batch: int, world: int, rank: int) -> torch.Tensor:
    """x: [n_samples * batch, ...] with flat index s * batch + b (what reparam produces,
    lib/models.py:23).  Returns this rank's [n_samples * local_batch, ...] with the same
    (s-major) ordering, so sample-axis reductions stay local."""
    lo, hi = batch_bounds(batch, world, rank)
    v = x.reshape(n_samples, batch, *x.shape[1:])
    return v[:, lo:hi].reshape(n_samples * (hi - lo), *x.shape[1:])


def shard_time_major(y: torch.Tensor, n_samples: int, batch: int, world: int, rank: int) -> torch.Tensor:
    """Same split for solver outputs / cotangents [T, n_samples * batch, ...]."""
    lo, hi = batch_bounds(batch, world, rank)
    v = y.reshape(y.shape[0], n_samples, batch, *y.shape[2:])
    return v[:, :, lo:hi].reshape(y.shape[0], n_samples * (hi - lo), *y.shape[2:])


def gather_time_major(y_local: torch.Tensor, n_samples: int, batch: int, group=None) -> torch.Tensor:
    """Inverse of shard_time_major (all-gather; for validation / evaluation only, not on the hot path)."""
    world = dist.get_world_size(group)
    T = y_local.shape[0]
    parts = []
    for r in range(world):
        lo, hi = batch_bounds(batch, world, r)
        parts.append(torch.empty((T, n_samples, hi - lo, *y_local.shape[2:]), dtype=y_local.dtype, device=y_local.device))
    mine = y_local.reshape(T, n_samples, -1, *y_local.shape[2:]).contiguous()
    dist.all_gather(parts, mine, group=group) if len({p.shape for p in parts}) == 1 else _uneven_all_gather(parts, mine, group)
    return torch.cat(parts, dim=2).reshape(T, n_samples * batch, *y_local.shape[2:])


def _uneven_all_gather(parts, mine, group):
    rank = dist.get_rank(group)
    for r, p in enumerate(parts):
        if r == rank:
            p.copy_(mine)
        dist.broadcast(p, src=dist.get_global_rank(group, r) if group is not None else r, group=group)


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x.  Backward is the identity: every rank forms the SAME global loss
    term from y, so dL/dx_local = dL/dy; the per-rank parameter gradients that follow are partial
    sums which all_reduce_grads() adds up exactly once."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


def all_reduce_side(side: torch.Tensor, group=None) -> torch.Tensor:
    """Differentiable all-reduce of the side-channel sums [.., 8] (fp64)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return side
    return _AllReduceSum.apply(side, group)


def all_reduce_grads(params: Iterable[torch.nn.Parameter], group=None, flat_buffer: Optional[torch.Tensor] = None):
    """Sum the ODE-func gradients over ranks in ONE message (27-294 KB: latency-bound on NVLink)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps]) if flat_buffer is None else flat_buffer
    if flat_buffer is not None:
        off = 0
        for p in ps:
            flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
            off += p.numel()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in ps:
        p.grad.copy_(flat[off:off + p.numel()].reshape(p.grad.shape))
        off += p.numel()
