"""Tube-Link (MMDetection) flavour of the temporal layers.

`MaXTron_Tube-Link/mmdet/models/plugins/msdeformattn_pixel_decoder.py` carries a third copy of the within-clip
classes -- TrajectoryAttention :652-708, TemporalEncoder :711-727, TemporalAxialTrajectoryAttentionLayer :730-791 --
with the same math and the same state-dict leaf names as the Video-kMaX copy, but they return the features only
(no attention maps) and `num_frames` defaults to 5.  These wrappers keep those signatures; the kernels are shared.
`temporal_branch` restates the temporal part of `MultiScaleDeformableAxialTrajectoryAttention.forward` (:616-632).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn
from torch import Tensor

from . import modules as _vk


class TrajectoryAttention(_vk.TrajectoryAttention):
    def forward(self, query, key, value, num_frames=5):
        return super().forward(query, key, value, num_frames)[0]


class TemporalAxialTrajectoryAttentionLayer(_vk.TemporalAxialTrajectoryAttentionLayer):
    def forward(self, src: Tensor, pos: Tensor):
        return self._run(src, pos)


class TemporalEncoder(nn.Module):
    """TL :711-727: only the axial layer type exists in the Tube-Link copy."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.0, attn_drop=0.0, activation="relu", n_heads=8, num_temporal_layer=2):
        super().__init__()
        self.temporal_layers = nn.ModuleList([TemporalAxialTrajectoryAttentionLayer(d_model, d_ffn, dropout, attn_drop, activation, n_heads)
                                              for _ in range(num_temporal_layer)])

    def forward(self, src: Tensor, pos: Tensor):
        for layer in self.temporal_layers:
            src = layer(src, pos)
        return src


def temporal_branch(level_feats: Sequence[Tensor], pos3d: Sequence[Tensor], temporal_layer: nn.Module, gamma: Tensor,
                    num_temporal_levels: int) -> List[Tensor]:
    """`f + gamma * temporal_layer(f, pos3d[i])` for the first `num_temporal_levels` levels (TL :620-627); the remaining
    levels pass through.  `gamma` is the learnable per-channel skip scale (init 1e-6, :485-486)."""
    out = []
    for i, f in enumerate(level_feats):
        if i < num_temporal_levels:
            out.append(torch.addcmul(f, gamma.to(f.dtype), temporal_layer(f, pos3d[i]).to(f.dtype)))
        else:
            out.append(f)
    return out
