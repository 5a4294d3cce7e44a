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


_level_streams = {}


def run_levels_concurrent(temporal_layer, feats: Sequence[Tensor], pos_3d: Sequence[Tensor], clip_chunks: int = 1):
    if clip_chunks > 1:
        return _run_chunked(temporal_layer, feats, pos_3d, clip_chunks)
    return _run_levels(temporal_layer, feats, pos_3d)


def _run_chunked(temporal_layer, feats: Sequence[Tensor], pos_3d: Sequence[Tensor], clip_chunks: int):
    """Clips are independent too: split every level's batch into `clip_chunks` groups and give each (level, group) its own
    stream, so that memory-bound and tensor-bound kernels of different groups overlap on the GPU."""
    tasks_f, tasks_p, owner = [], [], []
    for li, (f, p) in enumerate(zip(feats, pos_3d)):
        B, T = p.shape[0], p.shape[1]
        n = min(clip_chunks, B)
        bounds = [(B * k) // n for k in range(n + 1)]
        for k in range(n):
            b0, b1 = bounds[k], bounds[k + 1]
            tasks_f.append(f[b0 * T:b1 * T])
            tasks_p.append(p[b0:b1])
            owner.append(li)
    outs = _run_levels(temporal_layer, tasks_f, tasks_p)
    merged = []
    for li in range(len(feats)):
        parts = [o for o, ow in zip(outs, owner) if ow == li]
        merged.append((torch.cat([q[0] for q in parts], 0), parts[-1][1], parts[-1][2]))
    return merged


def _run_levels(temporal_layer, feats: Sequence[Tensor], pos_3d: Sequence[Tensor]):
    """Run the SAME temporal layer on several pyramid levels (WC/msdeformattn.py:261-263), one CUDA stream per level.

    The levels are independent, and each level is a chain of persistent one-CTA-per-SM kernels whose last wave leaves SMs
    idle (res5 at 32 clips: 221 tiles on 148 SMs); with one stream per level the block scheduler fills those SMs with the
    other level's CTAs.  Returns a list of (features, h_map, w_map) tuples in level order; the calling stream waits for all."""
    if len(feats) == 1:
        return [temporal_layer(src=feats[0], pos=pos_3d[0])]
    dev = feats[0].device
    cur = torch.cuda.current_stream(dev)
    key = (dev.index, len(feats))
    if key not in _level_streams:
        _level_streams[key] = [torch.cuda.Stream(dev) for _ in range(len(feats) - 1)]
    streams = [cur] + _level_streams[key]
    start = torch.cuda.Event()
    start.record(cur)
    outs = []
    for st, f, p in zip(streams, feats, pos_3d):
        if st is not cur:
            st.wait_event(start)
        with torch.cuda.stream(st):
            o = temporal_layer(src=f, pos=p)
            outs.append(o)
            if st is not cur:
                for t in o:
                    if isinstance(t, torch.Tensor):
                        t.record_stream(cur)
    for st in streams[1:]:
        cur.wait_stream(st)
    return outs


def run_temporal_levels(temporal_layer, memory: Tensor, spatial_shapes: Sequence[Tuple[int, int]], pos_3d: Sequence[Tensor],
                        num_temporal_levels: int):
    """memory [B*T, sum(H_l*W_l), C] -> same shape; returns (memory, height_traj_attn, width_traj_attn) like the reference."""
    sizes = [int(h) * int(w) for h, w in spatial_shapes]
    parts = list(torch.split(memory, sizes, dim=1))
    h_attn = w_attn = None
    for i in range(min(num_temporal_levels, len(parts))):
        parts[i], h_attn, w_attn = temporal_layer(src=parts[i].contiguous(), pos=pos_3d[i])
    return torch.cat(parts, dim=1), h_attn, w_attn


class WithinClipEncoder(torch.nn.Module):
    """Drop-in for `MSDeformAttnTransformerEncoder` (WC/msdeformattn.py:217-273): `num_stages` x [MSDeformAttn spatial layer on all
    levels, then the SAME-stage TemporalEncoder on the first `num_temporal_levels` levels].  Same sub-module names
    (`spatial_layers.{i}`, `temporal_layers.{i}`), forward signature and return tuple; unpadded feature maps only."""

    def __init__(self, spatial_layer, num_stages, transformer_num_spatial_feature_levels, transformer_num_temporal_feature_levels=0,
                 temporal_layer=None):
        super().__init__()
        import copy
        self.spatial_layers = torch.nn.ModuleList([copy.deepcopy(spatial_layer) for _ in range(num_stages)])
        self.transformer_num_spatial_feature_levels = transformer_num_spatial_feature_levels
        self.transformer_num_temporal_feature_levels = transformer_num_temporal_feature_levels
        if transformer_num_temporal_feature_levels > 0:
            self.temporal_layers = torch.nn.ModuleList([copy.deepcopy(temporal_layer) for _ in range(num_stages)])

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        from . import msda
        if valid_ratios is not None and not bool((valid_ratios == 1).all()):
            raise NotImplementedError("axial_vs_b200: padded feature maps (valid_ratios != 1) are not supported")
        shapes = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes) else spatial_shapes)]
        n = valid_ratios.shape[0] if valid_ratios is not None else 1
        return msda.reference_points(shapes, n, device)

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos, padding_mask, pos_3d=None):
        output = src
        shapes = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes) else spatial_shapes)]
        reference_points = self.get_reference_points(shapes, valid_ratios, src.device)
        if reference_points.shape[0] != src.shape[0]:
            reference_points = reference_points[:1].expand(src.shape[0], -1, -1, -1).contiguous()
        h_attn = w_attn = None
        for i, spatial_layer in enumerate(self.spatial_layers):
            output = spatial_layer(output, pos, reference_points, shapes, level_start_index, padding_mask)
            if self.transformer_num_temporal_feature_levels > 0:
                output, h_attn, w_attn = run_temporal_levels(self.temporal_layers[i], output, shapes, pos_3d,
                                                             self.transformer_num_temporal_feature_levels)
        return output, h_attn, w_attn
