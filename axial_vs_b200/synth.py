"""Deterministic synthetic weights and inputs for the hot path (tests, smoke, bench).

Follows SURVEY.md section 8(d) "Synthetic inputs": xavier_uniform_ on every dim>1 parameter (what the
reference's `_reset_parameters` does, WC/msdeformattn.py:70-73 and CC:148-151), nn.Linear-style
uniform biases, `src ~ N(0,1)`.  LayerNorm affine parameters are perturbed away from (1, 0) so that
the affine part of the norms is actually exercised by the parity tests.

All generators are seeded CPU `torch.Generator`s, so the same call reproduces the same tensors in the
authoring container (where the golden fixtures are made) and on the GPU box.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

Params = Dict[str, torch.Tensor]


def _xavier(g: torch.Generator, *shape: int) -> torch.Tensor:
    fan_out, fan_in = shape[0], shape[1]
    rf = 1
    for s in shape[2:]:
        rf *= s
    bound = math.sqrt(6.0 / ((fan_in + fan_out) * rf))
    return (torch.rand(*shape, generator=g) * 2 - 1) * bound


def _bias(g: torch.Generator, n: int, fan_in: int) -> torch.Tensor:
    b = 1.0 / math.sqrt(fan_in)
    return (torch.rand(n, generator=g) * 2 - 1) * b


def _ln(g: torch.Generator, n: int, prefix: str, out: Params) -> None:
    out[prefix + ".weight"] = 1.0 + 0.1 * torch.randn(n, generator=g)
    out[prefix + ".bias"] = 0.1 * torch.randn(n, generator=g)


def _linear(g: torch.Generator, prefix: str, n_out: int, n_in: int, out: Params) -> None:
    out[prefix + ".weight"] = _xavier(g, n_out, n_in)
    out[prefix + ".bias"] = _bias(g, n_out, n_in)


def traj_attn_params(g: torch.Generator, prefix: str, C: int, out: Params, fused_qkv: bool = False) -> None:
    """Leaf names of `TrajectoryAttention` (WC/temporal_attention.py:27-33; CC:85-89 for fused qkv)."""
    if fused_qkv:
        _linear(g, prefix + "qkv", 3 * C, C, out)
    else:
        for nm in ("q", "k", "v"):
            _linear(g, prefix + nm, C, C, out)
    _linear(g, prefix + "proj_q", C, C, out)
    _linear(g, prefix + "proj_kv", 2 * C, C, out)
    _linear(g, prefix + "proj", C, C, out)


def axial_layer_params(seed: int, C: int = 256, d_ffn: int = 1024, axial: bool = True) -> Params:
    """State dict of one Temporal(Axial)TrajectoryAttentionLayer (WC/temporal_attention.py:159-175)."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    if axial:
        traj_attn_params(g, "height_attn.", C, p)
        traj_attn_params(g, "width_attn.", C, p)
    else:
        traj_attn_params(g, "temporal_attn.", C, p)
    _ln(g, C, "norm1", p)
    _linear(g, "linear1", d_ffn, C, p)
    _linear(g, "linear2", C, d_ffn, p)
    _ln(g, C, "norm2", p)
    return p


def encoder_params(seed: int, num_layers: int = 2, C: int = 256, d_ffn: int = 1024, axial: bool = True) -> Params:
    """State dict of a TemporalEncoder: 'temporal_layers.{i}.<leaf>' (WC/temporal_attention.py:85-88)."""
    out: Params = {}
    for i in range(num_layers):
        for k, v in axial_layer_params(seed * 1000 + i, C, d_ffn, axial).items():
            out[f"temporal_layers.{i}.{k}"] = v
    return out


def cross_clip_params(seed: int, num_layers: int, num_classes: int = 124, C: int = 256) -> Params:
    """State dict of CrossClipTrackingModule (CC:204-272) with non-trivial BN running statistics."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}

    def bn(prefix: str, n: int, weight: float = 1.0):
        p[prefix + ".weight"] = weight * (1.0 + 0.1 * torch.randn(n, generator=g))
        p[prefix + ".bias"] = 0.1 * torch.randn(n, generator=g)
        p[prefix + ".running_mean"] = 0.1 * torch.randn(n, generator=g)
        p[prefix + ".running_var"] = 1.0 + 0.2 * torch.rand(n, generator=g)
        p[prefix + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.long)

    for i in range(num_layers):
        pre = f"transformer_trajectory_self_attention_layers.{i}."
        traj_attn_params(g, pre + "self_attn.", C, p, fused_qkv=True)
        _ln(g, C, pre + "norm", p)
        pre = f"conv_short_aggregate_layers.{i}."
        for j in range(3):
            p[pre + f"_aspp_conv{j}.weight"] = _xavier(g, C, C, 3)
            p[pre + f"_aspp_conv{j}.bias"] = _bias(g, C, 3 * C)
        p[pre + "_proj_conv_bn_act.conv.weight"] = _xavier(g, C, 3 * C, 1)
        _ln(g, C, pre + "_proj_conv_bn_act.norm", p)
        _ln(g, C, f"conv_norms.{i}", p)
    for nm in ("_class_embedding_projection", "_mask_embedding_projection"):
        p[nm + ".conv.weight"] = _xavier(g, C, C, 1)
        bn(nm + ".norm", C)
    p["_predictor._transformer_mask_head.conv.weight"] = _xavier(g, 128, C, 1)
    bn("_predictor._transformer_mask_head.norm", 128)
    p["_predictor._transformer_class_head.conv.weight"] = 0.01 * torch.randn(num_classes + 1, C, 1, generator=g)
    p["_predictor._transformer_class_head.conv.bias"] = torch.zeros(num_classes + 1)
    p["_predictor._transformer_class_activation_head.conv.weight"] = 0.01 * torch.randn(1, C, 1, generator=g)
    p["_predictor._transformer_class_activation_head.conv.bias"] = torch.zeros(1)
    bn("_predictor._pixel_space_mask_batch_norm", 1, weight=0.1)
    return p


def randn(seed: int, *shape: int) -> torch.Tensor:
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def level_embed(seed: int, levels: int = 2, C: int = 256) -> torch.Tensor:
    """`level_embed_3d ~ N(0,1)` (WC/msdeformattn.py:79-80)."""
    return randn(seed, levels, C)


def checksum(p: Params) -> float:
    """Order-independent fingerprint of a parameter dict (stored in the golden fixtures)."""
    tot = 0.0
    for k in sorted(p):
        t = p[k].double()
        tot += float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64).reshape(t.shape).remainder(7.0)).sum())
    return tot


def proj_params(seed: int, c: int):
    """State dicts of an input projection (c -> 256) and an output projection (256 -> 2c): Conv2d 1x1 + GroupNorm(32)."""
    g = torch.Generator().manual_seed(seed)
    pin = {"0.weight": _xavier(g, 256, c, 1, 1), "0.bias": 0.1 * torch.randn(256, generator=g),
           "1.weight": 1.0 + 0.2 * torch.randn(256, generator=g), "1.bias": 0.1 * torch.randn(256, generator=g)}
    pout = {"0.weight": _xavier(g, 2 * c, 256, 1, 1), "0.bias": 0.1 * torch.randn(2 * c, generator=g),
            "1.weight": 1.0 + 0.2 * torch.randn(2 * c, generator=g), "1.bias": 0.1 * torch.randn(2 * c, generator=g)}
    return pin, pout


def msda_layer_params(seed: int, n_levels: int = 3, n_heads: int = 8, n_points: int = 4, C: int = 256, d_ffn: int = 1024) -> Params:
    """State dict of MSDeformAttnTransformerEncoderLayer (WC/msdeformattn.py:177-203) with NON-degenerate sampling heads
    (the reference's init zeroes the offset / attention weights, which would hide layout mistakes)."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    no = n_heads * n_levels * n_points
    p["self_attn.sampling_offsets.weight"] = 0.05 * torch.randn(2 * no, C, generator=g)
    p["self_attn.sampling_offsets.bias"] = 1.5 * torch.randn(2 * no, generator=g)
    p["self_attn.attention_weights.weight"] = 0.05 * torch.randn(no, C, generator=g)
    p["self_attn.attention_weights.bias"] = 0.1 * torch.randn(no, generator=g)
    _linear(g, "self_attn.value_proj", C, C, p)
    _linear(g, "self_attn.output_proj", C, C, p)
    _ln(g, C, "norm1", p)
    _linear(g, "linear1", d_ffn, C, p)
    _linear(g, "linear2", C, d_ffn, p)
    _ln(g, C, "norm2", p)
    return p


def within_clip_module_params(seed: int, channels, num_stages: int = 2, temporal_layers_per_stage: int = 1) -> Params:
    """State dict of the whole within-clip tracking module (MSDeformAttnPixelDecoder, WC/msdeformattn.py:293-402) for feature
    levels given top-down (`channels` = [res5, res4, res3] widths): projections, level embeddings, spatial and temporal layers."""
    p: Params = {}
    g = torch.Generator().manual_seed(seed)
    for i, c in enumerate(channels):
        p[f"input_proj.{i}.0.weight"] = _xavier(g, 256, c, 1, 1)
        p[f"input_proj.{i}.0.bias"] = 0.1 * torch.randn(256, generator=g)
        p[f"input_proj.{i}.1.weight"] = 1.0 + 0.2 * torch.randn(256, generator=g)
        p[f"input_proj.{i}.1.bias"] = 0.1 * torch.randn(256, generator=g)
        p[f"output_proj.{i}.0.weight"] = _xavier(g, c, 256, 1, 1)
        p[f"output_proj.{i}.0.bias"] = 0.1 * torch.randn(c, generator=g)
        p[f"output_proj.{i}.1.weight"] = 1.0 + 0.2 * torch.randn(c, generator=g)
        p[f"output_proj.{i}.1.bias"] = 0.1 * torch.randn(c, generator=g)
    p["transformer.level_embed_2d"] = torch.randn(len(channels), 256, generator=g)
    p["transformer.level_embed_3d"] = torch.randn(2, 256, generator=g)
    for i in range(num_stages):
        p.update({f"transformer.encoder.spatial_layers.{i}.{k}": v for k, v in msda_layer_params(seed + 1 + i, len(channels)).items()})
        p.update({f"transformer.encoder.temporal_layers.{i}.{k}": v
                  for k, v in encoder_params(seed + 11 + i, temporal_layers_per_stage).items()})
    return p


def panoptic_case(seed: int, N: int, C: int, T: int, H: int, W: int, cell: int = 4, emb: int = 128):
    """Inputs of `panoptic_mask_inference` (Vk/maxtron_deeplab/maxtron_wc_model.py:439): class logits [N, C+1], mask logits
    [N, T, H, W] made of piecewise-constant cells (so slots own coherent regions that overlap at their borders) plus pixel noise,
    and mask embeddings [N, emb]."""
    g = torch.Generator().manual_seed(seed)
    mask_cls = torch.randn(N, C + 1, generator=g) * 3.0
    mask_cls[:, -1] -= 2.0                                            # keep the void class from winning everywhere
    hc, wc = (H + cell - 1) // cell, (W + cell - 1) // cell
    coarse = torch.randn(N, T, hc, wc, generator=g) * 2.5
    fine = coarse.repeat_interleave(cell, dim=2).repeat_interleave(cell, dim=3)[:, :, :H, :W]
    mask_pred = (fine + torch.randn(N, T, H, W, generator=g) * 0.7).contiguous()
    mask_embedding = torch.randn(N, emb, generator=g)
    return mask_cls, mask_pred, mask_embedding


def panoptic_metadata(C: int, label_divisor: int = 1000):
    """A VIPSeg-like split: contiguous ids 0..C-1, every third class is a thing; dataset ids are the contiguous ids + 1."""
    thing = {c + 1: c for c in range(C) if c % 3 == 0}
    stuff = {c + 1: c for c in range(C) if c % 3 != 0}
    return thing, stuff, label_divisor


def kmax_axial_params(seed: int, in_planes: int, key_depth: int = 512, value_depth: int = 1024, heads: int = 8) -> Params:
    """State dict of the kMaX `AxialAttention` (Vk/kmax_deeplab/modeling/pixel_decoder/kmax_pixel_decoder.py:105-126): 1x1 conv, three
    relative-position embedding tables (trunc-normal-like, std 1), three batch norms with non-trivial running statistics."""
    g = torch.Generator().manual_seed(seed)
    n_qkv = 2 * key_depth + value_depth
    p: Params = {"qkv_transform.conv.weight": torch.randn(n_qkv, in_planes, 1, generator=g) * in_planes ** -0.5}
    for name, depth in (("query", key_depth // heads), ("key", key_depth // heads), ("value", value_depth // heads)):
        p[f"_{name}_rpe._embeddings.weight"] = torch.randn(2 * 255 - 1, depth, generator=g).clamp_(-2, 2)
    for name, ch in (("_batch_norm_qkv", n_qkv), ("_batch_norm_similarity", 3 * heads), ("_batch_norm_retrieved_output", 2 * value_depth)):
        p[name + ".weight"] = 1.0 + 0.2 * torch.randn(ch, generator=g)
        p[name + ".bias"] = 0.1 * torch.randn(ch, generator=g)
        p[name + ".running_mean"] = 0.1 * torch.randn(ch, generator=g)
        p[name + ".running_var"] = 0.5 + torch.rand(ch, generator=g)
        p[name + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.long)
    # similarity logits are sums over 64 channels of unit-variance products: keep their batch-norm scale small so the softmax is not one-hot
    p["_batch_norm_similarity.weight"] *= 0.15
    return p


def tl_decoder_layer_params(seed: int, C: int = 256, d_ffn: int = 2048) -> Params:
    """State dict of the Tube-Link `DetrTransformerDecoderLayer` (mmcv key names): two MultiheadAttention blocks, FFN, three LayerNorms."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    for i in range(2):
        p[f"attentions.{i}.attn.in_proj_weight"] = _xavier(g, 3 * C, C)
        p[f"attentions.{i}.attn.in_proj_bias"] = 0.1 * torch.randn(3 * C, generator=g)
        _linear(g, f"attentions.{i}.attn.out_proj", C, C, p)
    _linear(g, "ffns.0.layers.0.0", d_ffn, C, p)
    _linear(g, "ffns.0.layers.1", C, d_ffn, p)
    for i in range(3):
        _ln(g, C, f"norms.{i}", p)
    return p


def tl_decoder_case(seed: int, Nq: int, B: int, L: int, heads: int = 8, C: int = 256, blocked: float = 0.6):
    """query / key / positions / boolean attention mask of one decoder step; no query row is fully blocked (the head un-blocks such rows,
    TL cc head :877-879)."""
    g = torch.Generator().manual_seed(seed)
    query, qpos = torch.randn(Nq, B, C, generator=g), torch.randn(Nq, B, C, generator=g)
    key, kpos = torch.randn(L, B, C, generator=g), torch.randn(L, B, C, generator=g)
    mask = torch.rand(B, 1, Nq, L, generator=g).expand(B, heads, Nq, L) < blocked          # the head repeats one mask over the heads
    mask = mask.reshape(B * heads, Nq, L).clone()
    mask[mask.all(-1)] = False
    if Nq > 1:
        mask[0, 1] = True                                                                    # one row with a single visible key
        mask[0, 1, L // 2] = False
    return query, qpos, key, kpos, mask


def kmax_layer_params(seed: int, in_channel_pixel: int, num_classes: int) -> Params:
    """State dict of the clip-level `kMaXTransformerLayer` (reference key names; bottleneck 256, key depth 128, value depth 256, 8 heads) with
    non-trivial eval-mode batch-norm statistics and affines (the reference's norm_init = 0 would zero the residual updates)."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}

    def conv(name, o, c, k=1, groups=1, two_d=False, bias=False, std=None):
        shape = (o, c // groups, k, k) if two_d else (o, c // groups, k)
        std = math.sqrt(2.0 / c) if std is None else std
        p[name + ".conv.weight"] = std * torch.randn(*shape, generator=g)
        if bias:
            p[name + ".conv.bias"] = 0.1 * torch.randn(o, generator=g)

    def bn(name, c):
        p[name + ".weight"] = 1.0 + 0.2 * torch.randn(c, generator=g)
        p[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
        p[name + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        p[name + ".running_var"] = 1.0 + 0.3 * torch.rand(c, generator=g)

    def conv_bn(name, o, c, **kw):
        conv(name, o, c, **kw)
        bn(name + ".norm", o)

    conv_bn("_query_conv1_bn_act", 256, 256)
    conv_bn("_pixel_conv1_bn_act", 256, in_channel_pixel, two_d=True)
    conv_bn("_query_qkv_conv_bn", 512, 256, std=256 ** -0.5)
    conv_bn("_pixel_v_conv_bn", 256, 256, two_d=True, std=256 ** -0.5)
    bn("_query_self_attention._batch_norm_similarity", 8)
    bn("_query_self_attention._batch_norm_retrieved_value", 256)
    conv_bn("_query_conv3_bn", 256, 256)
    conv_bn("_query_ffn_conv1_bn_act", 2048, 256)
    conv_bn("_query_ffn_conv2_bn", 256, 2048)
    conv_bn("_predictor._pixel_space_head_conv0bnact", 256, 256, k=5, groups=256, two_d=True, std=0.2)
    conv_bn("_predictor._pixel_space_head_conv1bnact", 256, 256, two_d=True)
    conv_bn("_predictor._pixel_space_head_last_convbn", 128, 256, two_d=True, bias=True, std=0.2)
    conv_bn("_predictor._transformer_mask_head", 128, 256)
    conv("_predictor._transformer_class_head", num_classes, 256, bias=True, std=0.05)
    bn("_predictor._pixel_space_mask_batch_norm", 1)
    bn("_kmeans_query_batch_norm_retrieved_value", 256)
    conv_bn("_kmeans_query_conv3_bn", 256, 256)
    return p
