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
                        num_temporal_levels: int, inplace: bool = False):
    """memory [B*T, sum(H_l*W_l), C] -> same shape; returns (memory, height_traj_attn, width_traj_attn) like the reference.

    The temporal levels are independent (WC/msdeformattn.py:261-263) and run on one CUDA stream each.  `inplace=True` (the caller owns
    `memory`, e.g. the fresh output of the spatial layer): the results are written back into the levels' slices of `memory` instead of
    re-concatenating every level -- the levels without a temporal layer (res3 = 75 % of the tokens) are then never copied."""
    sizes = [int(h) * int(w) for h, w in spatial_shapes]
    n = min(num_temporal_levels, len(sizes))
    if n == 0:
        return memory, None, None
    parts = list(torch.split(memory, sizes, dim=1))
    if memory.is_cuda:
        outs = run_levels_concurrent(temporal_layer, [parts[i].contiguous() for i in range(n)], list(pos_3d[:n]))
    else:
        outs = [temporal_layer(src=parts[i].contiguous(), pos=pos_3d[i]) for i in range(n)]
    h_attn, w_attn = outs[-1][1], outs[-1][2]
    if inplace and memory.is_contiguous() and not memory.requires_grad:
        for i in range(n):
            parts[i].copy_(outs[i][0])
        return memory, h_attn, w_attn
    for i in range(n):
        parts[i] = outs[i][0]
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
        self._ref_cache = {}

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
        if valid_ratios is not None and not bool((valid_ratios == 1).all()):
            raise NotImplementedError("axial_vs_b200: padded feature maps (valid_ratios != 1) are not supported")
        key = (tuple(shapes), str(src.device))
        if key not in self._ref_cache:                                              # identical for every image: one broadcast row set, built once per pyramid
            self._ref_cache = {key: self.get_reference_points(shapes, None, src.device)[:1].contiguous()}
        reference_points = self._ref_cache[key]
        h_attn = w_attn = None
        for i, spatial_layer in enumerate(self.spatial_layers):
            output = spatial_layer(output, pos, reference_points, shapes, level_start_index, padding_mask)
            if self.transformer_num_temporal_feature_levels > 0:
                output, h_attn, w_attn = run_temporal_levels(self.temporal_layers[i], output, shapes, pos_3d,
                                                             self.transformer_num_temporal_feature_levels, inplace=True)   # `output` is the spatial layer's own result
        return output, h_attn, w_attn


class _EncoderOnly(torch.nn.Module):
    """Parameter layout of `MSDeformAttnTransformerEncoderOnly` (WC/msdeformattn.py:30-80): level embeddings + `encoder`."""

    def __init__(self, d_model, nhead, num_stages, num_spatial_layers, num_temporal_layers, temporal_attn_type, dim_feedforward, dropout,
                 attn_drop, num_spatial_feature_levels, num_temporal_feature_levels):
        super().__init__()
        from . import modules, msda
        if not (num_spatial_layers > 0 and num_temporal_layers > 0 and num_spatial_layers == num_stages):
            raise NotImplementedError("axial_vs_b200.WithinClipTrackingModule: spatial + temporal layers per stage (every shipped config)")
        spatial = msda.MSDeformAttnTransformerEncoderLayer(d_model, dim_feedforward, dropout, "relu", num_spatial_feature_levels, nhead, 4)
        temporal = modules.TemporalEncoder(d_model, dim_feedforward, dropout, attn_drop, "relu", nhead, temporal_attn_type,
                                           num_temporal_layers // num_stages)
        self.encoder = WithinClipEncoder(spatial, num_spatial_layers, num_spatial_feature_levels, num_temporal_feature_levels, temporal)
        self.level_embed_2d = torch.nn.Parameter(torch.randn(num_spatial_feature_levels, d_model))
        self.level_embed_3d = torch.nn.Parameter(torch.randn(num_temporal_feature_levels, d_model))


class WithinClipTrackingModule(torch.nn.Module):
    """Drop-in for `MSDeformAttnPixelDecoder` (WC/msdeformattn.py:293-435), the within-clip tracking module: input projections,
    2-D / 3-D positional terms, `num_stages` x [MSDeformAttn spatial layer + trajectory-attention temporal layers], output
    projections.  Same constructor keywords (`input_shape` maps a feature name to an object with `.channels` / `.stride`), same
    state-dict keys (`input_proj.{i}.{0,1}.*`, `output_proj.{i}.{0,1}.*`, `transformer.level_embed_{2d,3d}`,
    `transformer.encoder.{spatial,temporal}_layers.*`) and the same `forward_features(features) -> (out, h_attn, w_attn)`.
    Differences inside: the positional tables are built once per shape and shared by all frames, tokens stay token-major between
    the projections and the encoder, and no NCHW <-> token copies are made."""

    def __init__(self, input_shape, *, transformer_dropout, transformer_attn_drop, transformer_nheads, transformer_dim_feedforward,
                 transformer_num_stages, transformer_spatial_layers, transformer_temporal_layers, transformer_temporal_attn_type, conv_dims,
                 transformer_spatial_in_features, transformer_temporal_in_features, num_clip_frames, cross_clip_training):
        super().__init__()
        from .pos import PositionEmbeddingSine
        from .projections import InputProjection, OutputProjection
        self.transformer_temporal_layers, self.num_clip_frames, self.cross_clip_training = transformer_temporal_layers, num_clip_frames, cross_clip_training
        spatial = sorted(((k, v) for k, v in input_shape.items() if k in transformer_spatial_in_features), key=lambda kv: kv[1].stride)
        temporal = sorted(((k, v) for k, v in input_shape.items() if k in transformer_temporal_in_features), key=lambda kv: kv[1].stride)
        self.transformer_spatial_in_features = [k for k, _ in spatial]           # "res3" .. "res5"
        self.transformer_temporal_in_features = [k for k, _ in temporal]
        chans = [v.channels for _, v in spatial]
        self.transformer_num_spatial_feature_levels = len(spatial)
        self.transformer_num_temporal_feature_levels = len(temporal)
        self.input_proj = torch.nn.ModuleList([InputProjection(c, conv_dims) for c in chans[::-1]])      # low resolution first
        self.output_proj = torch.nn.ModuleList([OutputProjection(c, conv_dims) for c in chans[::-1]])
        self.transformer = _EncoderOnly(conv_dims, transformer_nheads, transformer_num_stages, transformer_spatial_layers,
                                        transformer_temporal_layers, transformer_temporal_attn_type, transformer_dim_feedforward,
                                        transformer_dropout, transformer_attn_drop, len(spatial), len(temporal))
        self._pos2d_cache = (None, None)
        self.pe_layer = PositionEmbeddingSine(conv_dims // 2, normalize=True)
        self.pe_layer_3d = PositionEmbeddingSine3D(conv_dims // 2, normalize=True)

    @torch.no_grad()
    def forward_features(self, features):
        names = self.transformer_spatial_in_features[::-1]                        # top-down: res5, res4, res3
        BT = features[names[0]].shape[0]
        B = BT // self.num_clip_frames if (self.training or self.cross_clip_training) else 1
        T = BT // B
        shapes = [(int(features[f].shape[2]), int(features[f].shape[3])) for f in names]
        x0 = features[names[0]]
        # the multi-level token tensor [BT, Len, 256] is allocated once and every input projection writes its level's slice
        # (the reference concatenates the projected levels, WC/msdeformattn.py:106)
        src = torch.empty(BT, sum(h * w for h, w in shapes), 256, dtype=torch.float32, device=x0.device)
        pos3d, start = [], 0
        for i, f in enumerate(names):
            H, W = shapes[i]
            self.input_proj[i](features[f], out=src[:, start:start + H * W])
            start += H * W
            if f in self.transformer_temporal_in_features:
                pos3d.append(self.pe_layer_3d.table(B, T, H, W, x0.device, self.transformer.level_embed_3d[len(pos3d)]))
        if x0.dtype != torch.float32:
            src = src.to(x0.dtype)
        y, h_attn, w_attn = self.transformer.encoder(src, shapes, None, None, self._pos2d(shapes, x0.device), None, pos3d)
        out = {}
        start = 0
        for i, (H, W) in enumerate(shapes):
            out[names[i]] = self.output_proj[i](y[:, start:start + H * W], H, W)   # the level's slice is read in place
            start += H * W
        return out, h_attn, w_attn

    def _pos2d(self, shapes, device) -> Tensor:
        """[1, Len, 256] = per level PositionEmbeddingSine table + level_embed_2d (WC/msdeformattn.py:103-105), broadcast over the frames;
        rebuilt only when the pyramid shape or the level embedding changes."""
        le = self.transformer.level_embed_2d
        key = (tuple(shapes), str(device), le.data_ptr(), le._version)
        if self._pos2d_cache[0] != key:
            pos = torch.cat([self.pe_layer.table(H, W, device) + le[i].detach().float() for i, (H, W) in enumerate(shapes)], dim=0)
            self._pos2d_cache = (key, pos[None].contiguous())
        return self._pos2d_cache[1]
