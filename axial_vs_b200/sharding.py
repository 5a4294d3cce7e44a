"""Clip / video sharding across the GPUs of one box.

Within-clip module: clips are independent (`for idx in range(length)`, Vk/maxtron_deeplab/maxtron_wc_model.py:292-304),
so clip indices are partitioned contiguously over ranks with replicated weights and NO collective inside the hot
path.  NCCL is used only to all-gather per-clip outputs after the path (BASELINE.json north_star).
Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition: item i -> rank i*world // n_items; returns [start, stop) of `rank`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    start = (rank * n_items + world - 1) // world
    stop = ((rank + 1) * n_items + world - 1) // world
    return start, stop


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def gather_clip_outputs(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather per-clip outputs [n_local, ...] into [n_items, ...] in clip order on every rank (ragged shards ok)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_items, world)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError(f"rank {dist.get_rank(group)} holds {local.shape[0]} clips, expected {sizes[dist.get_rank(group)]}")
    mx = max(sizes)
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)
