"""Batch sharding for the sampling path: every image's 2n-step trajectory is independent, so rank r simply
takes images [lo, hi) of the batch; the only collective is one final all-gather of the outputs (new in this build —
the reference writes PNGs per rank instead, utils_Img2Img.py:390-400, and shards with accelerate, :316-317).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi): the first `total % world_size` ranks take one extra item
    (same split as the reference's `split` helper, utils_misc.py:63-71)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} out of range for world size {world_size}")
    base, extra = divmod(total, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_sizes(total: int, world_size: int) -> List[int]:
    return [shard_range(total, r, world_size)[1] - shard_range(total, r, world_size)[0] for r in range(world_size)]


def average_gradients(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Data-parallel gradient averaging of the training step: ONE all-reduce over the flat gradient vector, in
    place (what DDP does in buckets for the reference, train.py:57-61 through accelerate).  NCCL on the GPU box, gloo in the CPU tests."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / dist.get_world_size(group))
    return flat


def gather_outputs(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather the per-rank outputs (ragged along dim 0) into the full batch, in rank order.
    NCCL over NVLink on the GPU box, gloo in the CPU tests."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], dim=0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
