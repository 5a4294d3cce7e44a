"""Level plumbing of the within-clip tracking module around the temporal layers.

`MSDeformAttnTransformerEncoder.forward` (WC/msdeformattn.py:244-266) splits the concatenated multi-level memory
`[B*T, sum(H_l*W_l), C]` by level, runs the SAME TemporalEncoder on the first `num_temporal_levels` levels (res5, then
res4) and re-concatenates with the untouched remaining levels.  `run_temporal_levels` is that step with the B200 layers;
`level_positions` builds the `pos_3d` list the layers consume (table + level_embed_3d, WC/msdeformattn.py:112-115,419).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
from torch import Tensor

from .pos import PositionEmbeddingSine3D


def level_positions(pe: PositionEmbeddingSine3D, level_embed_3d: Tensor, B: int, T: int, shapes: Sequence[Tuple[int, int]], device) -> List[Tensor]:
    """pos_3d[i] = PositionEmbeddingSine3D table of level i (channels-last [B,T,H,W,C]) + level_embed_3d[i]."""
    return [pe.table(B, T, H, W, device, level_embed_3d[i]) for i, (H, W) in enumerate(shapes)]


def run_temporal_levels(temporal_layer, memory: Tensor, spatial_shapes: Sequence[Tuple[int, int]], pos_3d: Sequence[Tensor],
                        num_temporal_levels: int):
    """memory [B*T, sum(H_l*W_l), C] -> same shape; returns (memory, height_traj_attn, width_traj_attn) like the reference."""
    sizes = [int(h) * int(w) for h, w in spatial_shapes]
    parts = list(torch.split(memory, sizes, dim=1))
    h_attn = w_attn = None
    for i in range(min(num_temporal_levels, len(parts))):
        parts[i], h_attn, w_attn = temporal_layer(src=parts[i].contiguous(), pos=pos_3d[i])
    return torch.cat(parts, dim=1), h_attn, w_attn
