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


def cross_clip_sharded(refine_fn, mask_fn, clip_query_local: torch.Tensor, pixel_local: torch.Tensor, n_clips: int,
                       group=None, gather=None):
    """Cross-clip tracking module over a clip-sharded video.

    Inside one video the module needs every clip's queries, so there is ONE exchange step: the clip queries
    (`cluster_centers`, [1, Q, T_local, C] per rank) are all-gathered; the cross-clip layers then run REDUNDANTLY on every
    rank (cheaper than a second exchange); each rank computes mask logits only for its own clips (the mask einsum is
    per clip, CC:62-67) and the low-resolution mask logits are all-gathered at the end.

    refine_fn(clip_query_full [1, Q, T, C]) -> (class_logits [1, Q, K+1], mask_kernels [T*Q, ...] rows (t, q))
    mask_fn(mask_kernels_local [T_local*Q, ...], pixel_local, T_local) -> mask logits of the local clips [Q, T_local, ...]
    gather(x [n_local, ...], n_items) -> [n_items, ...]  (default: `gather_clip_outputs` over `group`)
    Returns (class_logits, mask_logits [Q, T, ...]) on every rank.
    """
    if gather is None:
        gather = lambda x, n: gather_clip_outputs(x, n, group)
    _, Q, T_local, C = clip_query_local.shape
    cq_full = gather(clip_query_local[0].permute(1, 0, 2).contiguous(), n_clips)          # [T, Q, C] in clip order
    cls, mk = refine_fn(cq_full.permute(1, 0, 2).unsqueeze(0).contiguous())               # redundant on every rank
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    t0, t1 = shard_range(n_clips, rank, world) if world > 1 else (0, T_local)
    if t1 - t0 != T_local:
        raise ValueError(f"rank {rank} holds {T_local} clips, the contiguous partition of {n_clips} expects {t1 - t0}")
    ml_local = mask_fn(mk[t0 * Q:t1 * Q], pixel_local, T_local)                           # [Q, T_local, ...]
    ml = gather(ml_local.transpose(0, 1).contiguous(), n_clips).transpose(0, 1).contiguous()
    return cls, ml
