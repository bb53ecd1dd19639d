"""Host batch scheduler: shard independent MPC instances over the GPUs of one box.

Instances are independent (own x0 / yref / p / iterate), so the solve path needs no
collective: rank g owns the contiguous slice ``shard_range(B, g, G)`` and keeps its
warm-start iterate resident on its GPU.  The only exchange is the final gather of
``u0`` / ``status`` (torch.distributed all_gather: NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first B % world ranks get one extra instance."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    q, r = divmod(B, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard(t: torch.Tensor | None, B: int, rank: int, world: int):
    """Slice the batch dimension of a per-instance tensor; shared tensors pass through.

    The solver's argument shapes decide which is which (``BlasterMPC._mode``): a 1-D tensor (yref[ny], p[25]) is
    shared by every instance whatever its length -- B == ny must not slice it -- and a tensor of two or more
    dimensions is per instance (x0[B,nx], yref[B,ny] / [B,N+1,ny], p[B,25] / [B,N,25]) and must lead with B."""
    if t is None or t.dim() <= 1:
        return t
    if t.shape[0] != B:
        raise ValueError(f"per-instance tensor of shape {tuple(t.shape)} does not lead with the global batch size {B}")
    lo, hi = shard_range(B, rank, world)
    return t[lo:hi]


def gather_batch(local: torch.Tensor, B: int, group=None) -> torch.Tensor:
    """All-gather variable-sized batch shards back into one [B, ...] tensor (same on every rank)."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    q, r = divmod(B, world)
    cap = q + (1 if r else 0)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for g in range(world):
        lo, hi = shard_range(B, g, world)
        parts.append(out[g * cap:g * cap + (hi - lo)])
    return torch.cat(parts, dim=0)


class ShardedSolve:
    """Run ``solve_fn`` on this rank's slice of a global batch and gather u0 / status.

    ``solve_fn(x0, yref, p) -> (u0, X, U, status)`` is normally ``BlasterMPC.solve`` of a
    solver created with ``batch = shard size``; the tests inject a CPU stand-in."""

    def __init__(self, solve_fn, B: int, group=None):
        self.solve_fn, self.B, self.group = solve_fn, B, group
        init = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if init else 0
        self.world = dist.get_world_size(group) if init else 1
        self.lo, self.hi = shard_range(B, self.rank, self.world)

    def solve(self, x0, yref, p=None, gather: bool = True):
        B, r, w = self.B, self.rank, self.world
        u0, X, U, status = self.solve_fn(shard(x0, B, r, w), shard(yref, B, r, w), shard(p, B, r, w))
        if not gather:
            return u0, status
        return gather_batch(u0, B, self.group), gather_batch(status, B, self.group)
