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


# ------------------------------------------------------------------------------------------------------------------
# Tube-Link mask decoder layer (row A11, TL half): DetrTransformerDecoderLayer with operation_order
# ('cross_attn', 'norm', 'self_attn', 'norm', 'ffn', 'norm'), TL/mmdet/models/utils/transformer.py:408-451, configured at
# TL/configs/video/**: MultiheadAttention(embed_dims=256, num_heads=8, batch_first=False), FFN(256 -> 2048 -> 256, ReLU, add_identity),
# post-norm; called once per decoder step at TL/models/video/tube_link_vis/mask2former_video_cc_head.py:883-894 with
# attn_masks = [attn_mask, None].
#
# mmcv-full 1.6.1 (MultiheadAttention, BaseTransformerLayer, FFN) is NOT vendored in the reference tree and not installed here, so its
# wrapper semantics are RESTATED, not pinned -- "parity unpinned" for the wrapper:
#   MultiheadAttention.forward(query, key, value, identity, query_pos, key_pos, attn_mask):
#       q = query + query_pos; k = key + key_pos; v = value; out = torch.nn.MultiheadAttention(q, k, v, attn_mask)[0]; return identity + out
#   BaseTransformerLayer (post-norm): 'cross_attn' uses (key, value, key_pos, attn_masks[0]); 'self_attn' uses key = value = query,
#       key_pos = query_pos, attn_masks[1]; 'norm' = LayerNorm; 'ffn' = x + Linear(ReLU(Linear(x))).
# The arithmetic core IS pinned: the oracle restatement is checked against torch.nn.MultiheadAttention / nn.LayerNorm / nn.Linear themselves
# (tests/test_oracle_golden.py::test_tl_decoder_layer_oracle_against_torch_mha).
# State-dict keys follow mmcv: attentions.{0,1}.attn.{in_proj_weight,in_proj_bias,out_proj.weight,out_proj.bias},
# ffns.0.layers.0.0.{weight,bias}, ffns.0.layers.1.{weight,bias}, norms.{0,1,2}.{weight,bias}.
# ------------------------------------------------------------------------------------------------------------------
class _TorchMHAParams(nn.Module):
    """Parameter holder with torch.nn.MultiheadAttention's names (mmcv keeps the torch module as `.attn`)."""

    def __init__(self, embed_dims: int):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dims, embed_dims))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dims))
        self.out_proj = nn.Linear(embed_dims, embed_dims)
        nn.init.xavier_uniform_(self.in_proj_weight)


class MultiheadAttention(nn.Module):
    """mmcv.cnn.bricks.transformer.MultiheadAttention (batch_first=False, no dropout) on the library's masked attention kernel."""

    def __init__(self, embed_dims: int = 256, num_heads: int = 8, attn_drop: float = 0.0, proj_drop: float = 0.0, dropout_layer=None,
                 batch_first: bool = False, **kwargs):
        super().__init__()
        if embed_dims != num_heads * 32:
            raise NotImplementedError("MultiheadAttention: built for 32 channels per head (embed_dims = 32 * num_heads)")
        if batch_first:
            raise NotImplementedError("MultiheadAttention: the Tube-Link decoder uses batch_first=False")
        self.embed_dims, self.num_heads = embed_dims, num_heads
        self.attn = _TorchMHAParams(embed_dims)
        self._cache = _PackedCache()
        self.register_load_state_dict_post_hook(_invalidate_hook)

    def _packed(self, dev):
        def build():
            E = self.embed_dims
            w, b = self.attn.in_proj_weight.detach().float(), self.attn.in_proj_bias.detach().float()
            return {"wq": ops.pack_weight(w[:E].contiguous()), "wk": ops.pack_weight(w[E:2 * E].contiguous()), "wv": ops.pack_weight(w[2 * E:].contiguous()),
                    "bq": b[:E].contiguous(), "bk": b[E:2 * E].contiguous(), "bv": b[2 * E:].contiguous(),
                    "wo": ops.pack_weight(self.attn.out_proj.weight.detach().float().contiguous()), "bo": self.attn.out_proj.bias.detach().float().contiguous()}
        return self._cache.get(self, dev, build)

    def project_kv(self, key: Tensor, value: Tensor, key_pos):
        """k = (key + key_pos) Wk^T + bk and v = value Wv^T + bv as fp32 rows (the decoder re-uses a level's memory in three layers, but every
        layer has its own weights).  The positional term goes through the GEMM's residual input: (key + key_pos) W^T = key W^T + key_pos W^T."""
        pk = self._packed(key.device)
        E = self.embed_dims
        rows = key.reshape(-1, E)
        kp = ops.linear(ops.cast_bf16(key_pos.reshape(-1, E)), pk["wk"], None, E, out_dtype=torch.float32) if key_pos is not None else None
        k = ops.linear(ops.cast_bf16(rows), pk["wk"], pk["bk"], E, out_dtype=torch.float32, resid=kp)
        v = ops.linear(ops.cast_bf16(value.reshape(-1, E)), pk["wv"], pk["bv"], E, out_dtype=torch.float32)
        return k.view(key.shape), v.view(value.shape)

    def forward(self, query: Tensor, key: Tensor = None, value: Tensor = None, identity: Tensor = None, query_pos: Tensor = None,
                key_pos: Tensor = None, attn_mask: Tensor = None, key_padding_mask: Tensor = None, **kwargs) -> Tensor:
        _require_inference(self, query)
        if key_padding_mask is not None:
            raise NotImplementedError("MultiheadAttention: key_padding_mask is None at both call sites of the Tube-Link decoder")
        key = query if key is None else key
        value = key if value is None else value
        identity = query if identity is None else identity
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        pk = self._packed(query.device)
        E = self.embed_dims
        query, identity = query.float().contiguous(), identity.float().contiguous()
        scale = (E // self.num_heads) ** -0.5 * 1.4426950408889634           # softmax scale and log2(e): the kernel works in the exp2 domain
        # axvs_linear computes (a W^T + bias) * scale + resid: the positional term enters pre-scaled
        qp = ops.linear(ops.cast_bf16(query_pos.float().reshape(-1, E)), pk["wq"], None, E, scale=scale, out_dtype=torch.float32) if query_pos is not None else None
        q = ops.linear(ops.cast_bf16(query.reshape(-1, E)), pk["wq"], pk["bq"], E, scale=scale, out_dtype=torch.float32, resid=qp).view(query.shape)
        k, v = self.project_kv(key.float().contiguous(), value.float().contiguous(), None if key_pos is None else key_pos.float().contiguous())
        if attn_mask is not None and attn_mask.dtype != torch.bool:
            raise NotImplementedError("MultiheadAttention: boolean attn_mask only (True = blocked), as the Tube-Link head builds it")
        o = ops.masked_mha(q, k, v, None if attn_mask is None else attn_mask.contiguous(), heads=self.num_heads, seq_first=True)
        return ops.linear(o.reshape(-1, E), pk["wo"], pk["bo"], E, out_dtype=torch.float32, resid=identity.reshape(-1, E)).view(query.shape)


class _FFN(nn.Module):
    """mmcv FFN: layers = Sequential(Sequential(Linear, ReLU, Dropout), Linear, Dropout), add_identity."""

    def __init__(self, embed_dims: int, feedforward_channels: int):
        super().__init__()
        self.layers = nn.Sequential(nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(0.0)),
                                    nn.Linear(feedforward_channels, embed_dims), nn.Dropout(0.0))


class DetrTransformerDecoderLayer(nn.Module):
    """Drop-in for the Tube-Link decoder layer (see the block comment above): forward(query, key, value, query_pos, key_pos, attn_masks)."""

    def __init__(self, embed_dims: int = 256, num_heads: int = 8, feedforward_channels: int = 2048,
                 operation_order=("cross_attn", "norm", "self_attn", "norm", "ffn", "norm"), **kwargs):
        super().__init__()
        if tuple(operation_order) != ("cross_attn", "norm", "self_attn", "norm", "ffn", "norm"):
            raise NotImplementedError("DetrTransformerDecoderLayer: the operation order of the shipped Tube-Link configs only")
        self.embed_dims, self.pre_norm = embed_dims, False
        self.attentions = nn.ModuleList([MultiheadAttention(embed_dims, num_heads), MultiheadAttention(embed_dims, num_heads)])
        self.ffns = nn.ModuleList([_FFN(embed_dims, feedforward_channels)])
        self.norms = nn.ModuleList([nn.LayerNorm(embed_dims) for _ in range(3)])
        self._cache = _PackedCache()
        self.register_load_state_dict_post_hook(_invalidate_hook)

    def _packed_ffn(self, dev):
        def build():
            l0, l1 = self.ffns[0].layers[0][0], self.ffns[0].layers[1]
            return {"w1": ops.pack_weight(l0.weight.detach().float().contiguous()), "b1": l0.bias.detach().float().contiguous(),
                    "w2": ops.pack_weight(l1.weight.detach().float().contiguous()), "b2": l1.bias.detach().float().contiguous()}
        return self._cache.get(self.ffns[0], dev, build)

    def _norm(self, i: int, x: Tensor) -> Tensor:
        n = self.norms[i]
        return ops.layernorm(x.reshape(-1, self.embed_dims), n.weight.detach().float(), n.bias.detach().float(), n.eps).view(x.shape)

    def forward(self, query: Tensor, key: Tensor = None, value: Tensor = None, query_pos: Tensor = None, key_pos: Tensor = None,
                attn_masks=None, query_key_padding_mask=None, key_padding_mask=None, **kwargs) -> Tensor:
        _require_inference(self, query)
        if query_key_padding_mask is not None or key_padding_mask is not None:
            raise NotImplementedError("DetrTransformerDecoderLayer: padding masks are None in the Tube-Link head")
        masks = [None, None] if attn_masks is None else list(attn_masks)
        E = self.embed_dims
        x = self._norm(0, self.attentions[0](query, key, value, None, query_pos=query_pos, key_pos=key_pos, attn_mask=masks[0]))
        x = self._norm(1, self.attentions[1](x, x, x, None, query_pos=query_pos, key_pos=query_pos, attn_mask=masks[1]))
        pk = self._packed_ffn(x.device)
        F_ = self.ffns[0].layers[0][0].out_features
        h = ops.linear(ops.cast_bf16(x.reshape(-1, E)), pk["w1"], pk["b1"], F_, relu=True)
        y = ops.linear(h, pk["w2"], pk["b2"], E, out_dtype=torch.float32, resid=x.reshape(-1, E)).view(x.shape)
        return self._norm(2, y)
