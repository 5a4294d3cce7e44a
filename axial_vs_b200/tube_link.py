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


# ------------------------------------------------------------------------------------------------------------------
# MultiScaleDeformableAxialTrajectoryAttention (TL :393-638) and the cross-clip head's prediction tail (TL cc head :761-797)
# ------------------------------------------------------------------------------------------------------------------
import math
from typing import Optional

from . import ops
from .modules import _PackedCache, _invalidate_hook, _require_inference

TL_PD = "MaXTron_Tube-Link/mmdet/models/plugins/msdeformattn_pixel_decoder.py"


class MultiScaleDeformableAxialTrajectoryAttention(nn.Module):
    """Drop-in for the Tube-Link `ATTENTION` plugin of the same name (TL_PD:393-638): same constructor arguments (the MMCV dict keys
    `embed_dims, num_heads, num_levels, num_temporal_levels, num_temporal_layers, num_temporal_dim, num_points, skip_connect, attn_drop,
    dropout, batch_first`), forward signature and state-dict keys (`sampling_offsets`, `attention_weights`, `value_proj`, `output_proj`,
    `temporal_layer.temporal_layers.{k}.*`, `gamma`).  forward = deformable sampling (:573-614) -> per temporal level
    `f + gamma * temporal_layer(f, query_pos3d[i])` (:616-627) -> concat -> `output_proj` (:632) -> `+ identity` (:638).
    Inference only; unpadded maps, 2-D reference points (what the Tube-Link encoder passes), value_proj_ratio = 1."""

    def __init__(self, embed_dims: int = 256, num_heads: int = 8, num_levels: int = 4, num_temporal_levels: int = 2, num_temporal_layers: int = 1,
                 num_temporal_dim: int = 1024, num_points: int = 4, im2col_step: int = 64, dropout: float = 0.1, skip_connect: bool = True,
                 attn_drop: float = 0.0, batch_first: bool = False, norm_cfg=None, init_cfg=None, value_proj_ratio: float = 1.0):
        super().__init__()
        if embed_dims != ops.C or num_heads != ops.HEADS or value_proj_ratio != 1.0:
            raise NotImplementedError("axial_vs_b200 kernels are specialised for embed_dims=256, num_heads=8, value_proj_ratio=1")
        if num_levels > 4 or num_levels * num_points > 16:
            raise NotImplementedError("axial_vs_b200: at most 4 levels and 16 level*point samples per head")
        self.embed_dims, self.num_heads, self.num_levels, self.num_points = embed_dims, num_heads, num_levels, num_points
        self.num_temporal_levels, self.num_temporal_layers = num_temporal_levels, num_temporal_layers
        self.skip_connect, self.attn_drop, self.batch_first, self.im2col_step = skip_connect, attn_drop, batch_first, im2col_step
        self.dropout = nn.Dropout(dropout)
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.temporal_layer = TemporalEncoder(embed_dims, num_temporal_dim, attn_drop=attn_drop, num_temporal_layer=num_temporal_layers)
        if skip_connect:
            self.gamma = nn.Parameter(1e-6 * torch.ones(embed_dims))
        self.init_weights()
        self._cache = _PackedCache()
        self.register_load_state_dict_post_hook(_invalidate_hook)

    def init_weights(self) -> None:
        """TL_PD:488-511: zero offset weights with compass-direction biases, zero attention weights, Xavier value / output projections."""
        ang = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid = torch.stack([ang.cos(), ang.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.num_heads, 1, 1, 2).repeat(1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.weight.zero_()
            self.sampling_offsets.bias.copy_(grid.view(-1))
            self.attention_weights.weight.zero_()
            self.attention_weights.bias.zero_()
            for lin in (self.value_proj, self.output_proj):
                nn.init.xavier_uniform_(lin.weight)
                lin.bias.zero_()

    def _packed(self, device):
        def build():
            so, aw, vp, op_ = self.sampling_offsets, self.attention_weights, self.value_proj, self.output_proj
            sampler = ops.pack_msda_sampler(so.weight, so.bias, aw.weight, aw.bias, vp.weight, vp.bias, self.num_levels, self.num_points)
            return sampler, ops.pack_weight_split(op_.weight), op_.bias.detach().float().contiguous()
        return self._cache.get(nn.ModuleList([self.sampling_offsets, self.attention_weights, self.value_proj, self.output_proj]), device, build)

    def forward(self, query: Tensor, key: Optional[Tensor] = None, value: Optional[Tensor] = None, identity: Optional[Tensor] = None,
                query_pos: Optional[Tensor] = None, query_pos3d: Optional[List[Tensor]] = None, key_padding_mask: Optional[Tensor] = None,
                reference_points: Optional[Tensor] = None, spatial_shapes=None, level_start_index=None, **kwargs) -> Tensor:
        _require_inference(self, query)
        if key_padding_mask is not None and bool(key_padding_mask.any()):
            raise NotImplementedError("axial_vs_b200: padded feature maps are not supported")
        if reference_points is None or reference_points.shape[-1] != 2:
            raise NotImplementedError("axial_vs_b200: 2-D reference points only (encoder use)")
        value = query if value is None else value
        identity = query if identity is None else identity
        if not self.batch_first:                                          # (num_query, bs, C) -> (bs, num_query, C)     TL_PD:564-567
            query, value = query.permute(1, 0, 2), value.permute(1, 0, 2)
            query_pos = None if query_pos is None else query_pos.permute(1, 0, 2)
        shapes = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes) else spatial_shapes)]
        sampler, w_out, b_out = self._packed(query.device)
        q32 = query.contiguous().float()
        sampled = ops.msda_sample_fwd(value.contiguous().float(), q32, None if query_pos is None else query_pos.contiguous().float(),
                                      reference_points.contiguous().float(), shapes, sampler, self.num_levels, self.num_points)   # bf16 [bs, nq, C]
        outs = list(torch.split(sampled.float(), [h * w for h, w in shapes], dim=1))                        # :618-620
        for i in range(self.num_temporal_levels):                                                            # :622-627
            f = outs[i].contiguous()
            t = self.temporal_layer(src=f, pos=query_pos3d[i])
            outs[i] = torch.addcmul(f, self.gamma.to(f.dtype), t) if self.skip_connect else t
        cat = torch.cat(outs, dim=1)
        bs, nq, c = cat.shape
        out = ops.linear_f32(cat.view(bs * nq, c).contiguous(), w_out, b_out, c, 0, True).view(bs, nq, c)    # output_proj :632
        if not self.batch_first:
            out = out.permute(1, 0, 2)
        return (out + identity.float()).to(identity.dtype)                                                   # dropout = identity in eval :638


class CCHeadPredictor(nn.Module):
    """The prediction tail of `Mask2FormerVideoCCHeadTube` (TL cc head :761-797): `forward_head_clips` + `pred_class`, with the reference's
    sub-module names (`post_norm`, `activation_proj`, `cls_embed`, `mask_embed.{0,2,4}`) so the head's checkpoint entries load unchanged.
    The mask embedding MLP and the query x pixel contraction run split-precision (fp32-grade), like the Video-kMaX tail."""

    def __init__(self, feat_channels: int = 256, out_channels: int = 256, num_classes: int = 40):
        super().__init__()
        if feat_channels != 256 or out_channels % 128:
            raise NotImplementedError("axial_vs_b200: feat_channels must be 256 and out_channels a multiple of 128")
        self.post_norm = nn.LayerNorm(feat_channels)
        self.activation_proj = nn.Linear(feat_channels, 1)
        self.cls_embed = nn.Linear(feat_channels, num_classes + 1)
        self.mask_embed = nn.Sequential(nn.Linear(feat_channels, feat_channels), nn.ReLU(inplace=True), nn.Linear(feat_channels, feat_channels),
                                        nn.ReLU(inplace=True), nn.Linear(feat_channels, out_channels))
        self.out_channels, self.num_classes = out_channels, num_classes
        self._cache = _PackedCache()
        self.register_load_state_dict_post_hook(_invalidate_hook)

    def _packed(self, device):
        def pad(lin):
            w, b = lin.weight.detach().float(), lin.bias.detach().float()
            n_pad = (w.shape[0] + 255) // 256 * 256
            wp, bp = w.new_zeros(n_pad, w.shape[1]), b.new_zeros(n_pad)
            wp[: w.shape[0]], bp[: w.shape[0]] = w, b
            return ops.pack_weight_split(wp), bp, n_pad
        return self._cache.get(self, device, lambda: [pad(self.mask_embed[0]), pad(self.mask_embed[2]), pad(self.mask_embed[4]), pad(self.cls_embed)])

    def forward_head_clips(self, decoder_out: Tensor, mask_feature: Tensor):
        """decoder_out [t, l, q, b, c]; mask_feature [b, T_frames, c_m, h, w] -> (tuple of l class logits [b, q, K+1], tuple of l mask
        logits [b, T_frames, q, h, w]), as the reference returns them (`unbind(0)`)."""
        _require_inference(self, decoder_out, mask_feature)
        t, l, q, b, c = decoder_out.shape
        if b != 1:
            raise NotImplementedError("axial_vs_b200: one video at a time at inference (as the reference's test path)")
        Tf, cm = mask_feature.shape[1], mask_feature.shape[2]
        fpc = Tf // t
        x = ops.layernorm(decoder_out.contiguous().float().view(-1, c), self.post_norm.weight.detach().float(), self.post_norm.bias.detach().float(),
                          self.post_norm.eps).view(t, l, q, b, c)                                            # :768
        x = x.permute(1, 3, 0, 2, 4).contiguous()                                                           # (l, b, t, q, c)  :769
        (w0, b0, n0), (w2, b2, n2), (w4, b4, n4), (wc, bc, nc) = self._packed(x.device)
        # pred_class (:783-797): softmax over the clips of activation_proj, weighted sum, cls_embed
        pooled = torch.stack([ops.cc_class_pool(x[i, 0].reshape(t * q, c).bfloat16().contiguous(), self.activation_proj.weight.detach().float().reshape(-1).contiguous(),
                                                float(self.activation_proj.bias.detach()), t, q) for i in range(l)], 0)           # [l, q, c] bf16
        cls = ops.linear_f32(pooled.float().view(l * q, c).contiguous(), wc, bc, nc, 0, True)[:, : self.num_classes + 1].reshape(l, b, q, -1)
        # mask_embed MLP (:772) and the per-clip einsum (:774-777)
        rows = x.view(l * b * t * q, c)
        me = ops.linear_f32(ops.linear_f32(ops.linear_f32(rows, w0, b0, n0, 1, True), w2, b2, n2, 1, True), w4, b4, n4, 0, True)   # [l*t*q, n4]
        me = me.view(l, t, q, n4)
        masks = []
        for i in range(l):
            pix = mask_feature[0].reshape(t, fpc, cm, -1).permute(0, 2, 1, 3).reshape(t, cm, -1).contiguous().float()   # per clip [cm, fpc*h*w]
            ml = ops.mask_einsum(pix, me[i].reshape(t * q, n4).contiguous(), t, q, pix.shape[-1], 1.0, 0.0)    # [q, t, fpc*h*w]
            masks.append(ml.view(q, t * fpc, *mask_feature.shape[3:]).permute(1, 0, 2, 3).unsqueeze(0))       # [b, T_frames, q, h, w]
        return tuple(cls.unbind(0)), tuple(masks)
