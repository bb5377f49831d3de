"""Batch sharding of multi-view samples across the GPUs of one box.

Samples are independent, so rank r simply owns a contiguous slice of the batch; all V views of a sample stay on one
GPU.  The only collective is an optional all-gather of the [B_local, 21, 3] poses (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world_size: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of the samples rank `rank` owns; ragged batches give the first `batch % world_size` ranks one more."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_inputs(x, bbox=None, intr=None, world_size: int = 1, rank: int = 0):
    lo, hi = shard_bounds(x.shape[0], world_size, rank)
    return (x[lo:hi], None if bbox is None else bbox[lo:hi], None if intr is None else intr[lo:hi])


def gather_poses(joints_local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """All-gather ragged per-rank poses [B_r, 21, 3] back into batch order [batch, 21, 3]."""
    if not dist.is_available() or not dist.is_initialized():
        return joints_local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    padded = joints_local.new_zeros((cap,) + tuple(joints_local.shape[1:]))
    padded[: joints_local.shape[0]] = joints_local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)
