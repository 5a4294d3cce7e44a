"""TEST INFRASTRUCTURE (checker only).  CPU restatement of the clip-level kMaX decoder layer, pixel and query side:
kMaXPredictor.forward (DEC:95-124) and kMaXTransformerLayer.forward (DEC:184-232), DEC =
MaXTron_Video-kMaX/maxtron_deeplab/modeling/transformer_decoder/maxtron_transformer_decoder.py, eval mode (SyncBatchNorm -> running
statistics, eps 1e-3; DropPath = identity).  Pinned on the unmodified reference module: tests/golden/kmax_layer_{a,b}.npz written by
oracle/make_golden_kmax_layer.py, checked in tests/test_oracle_golden.py::test_kmax_layer_oracle_golden (and live, when the reference tree
is mounted, in tests/test_oracle_vs_reference.py)."""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import traj_oracle as O

Params = Dict[str, torch.Tensor]


def conv_bn_2d_1x1(x, p: Params, prefix: str, norm: bool, act):
    """ConvBN(kernel_size=1, conv_type='2d') on [N, C, H, W] -- kmax_pixel_decoder.py:42-72."""
    N, C, H, W = x.shape
    q = dict(p)
    q[prefix + ".conv.weight"] = p[prefix + ".conv.weight"][:, :, :, 0]          # [O, C, 1, 1] -> the 1-D form conv_bn_1d expects
    return O.conv_bn_1d(x.reshape(N, C, H * W), q, prefix, "syncbn" if norm else None, act).reshape(N, -1, H, W)


def depthwise5_bn_gelu(x, p: Params, prefix: str):
    """ConvBN(C, C, kernel_size=5, groups=C, padding=2, bias=False, norm='syncbn', act='gelu') -- DEC:78-79."""
    N, C, H, W = x.shape
    w = p[prefix + ".conv.weight"].to(x.dtype)                      # [C, 1, 5, 5]
    xp = torch.zeros(N, C, H + 4, W + 4, dtype=x.dtype)
    xp[:, :, 2:2 + H, 2:2 + W] = x
    y = torch.zeros_like(x)
    for dy in range(5):
        for dx in range(5):
            y = y + xp[:, :, dy:dy + H, dx:dx + W] * w[None, :, 0, dy, dx, None, None]
    return O.gelu(O.batch_norm_eval(y.reshape(N, C, H * W), p, prefix + ".norm").reshape(N, C, H, W))


def predictor(mask_emb, class_emb, pixel_feature, p: Params, prefix: str = "_predictor."):
    f = depthwise5_bn_gelu(pixel_feature, p, prefix + "_pixel_space_head_conv0bnact")                     # :98
    f = conv_bn_2d_1x1(f, p, prefix + "_pixel_space_head_conv1bnact", True, "gelu")                       # :99
    f = conv_bn_2d_1x1(f, p, prefix + "_pixel_space_head_last_convbn", True, None)                        # :100
    f = f / f.norm(dim=1, keepdim=True).clamp_min(1e-12)                                                  # :101  F.normalize(p=2, dim=1)
    cls = O.conv_bn_1d(class_emb, p, prefix + "_transformer_class_head", None, None).permute(0, 2, 1)     # :103
    K = cls.shape[-1]
    bias = torch.zeros(K, dtype=cls.dtype)
    bias[-1] = math.log((K - 1) * 0.9 / (1 - 0.9))                                                        # :104 add_bias_towards_void
    cls = cls + bias
    mk = O.conv_bn_1d(mask_emb, p, prefix + "_transformer_mask_head", "syncbn", None)                     # :105
    logits = torch.einsum("bchw,bcn->bnhw", f, mk)                                                        # :106-107
    N, L, H, W = logits.shape
    logits = O.batch_norm_eval(logits.reshape(N, 1, L * H * W), p, prefix + "_pixel_space_mask_batch_norm").reshape(N, L, H, W)   # :110
    return {"class_logits": cls, "mask_logits": logits, "mask_embeddings": mk.permute(0, 2, 1), "pixel_feature": f}


def transformer_layer(pixel_feature, query_feature, p: Params, heads: int = 8, key_depth: int = 128, value_depth: int = 256, advanced: bool = False):
    N, C, TH, W = pixel_feature.shape
    L = query_feature.shape[2]
    pixel_space = conv_bn_2d_1x1(O.gelu(pixel_feature), p, "_pixel_conv1_bn_act", True, "gelu")           # :186
    query_space = O.conv_bn_1d(query_feature, p, "_query_conv1_bn_act", "syncbn", "gelu")                 # :187
    pixel_value = conv_bn_2d_1x1(pixel_space, p, "_pixel_v_conv_bn", True, None).reshape(N, value_depth, TH * W)   # :190-191
    pred = predictor(query_space, query_space, pixel_space, p)                                            # :193-194
    km = O.kmeans_update(pred["mask_logits"].flatten(2), pixel_value, advanced)                           # :196-208
    km = km[0] if isinstance(km, tuple) else km
    km = O.batch_norm_eval(km, p, "_kmeans_query_batch_norm_retrieved_value")                             # :209
    q = query_feature + O.conv_bn_1d(km, p, "_kmeans_query_conv3_bn", "syncbn", None)                     # :210-211
    qkv = O.conv_bn_1d(query_space, p, "_query_qkv_conv_bn", "syncbn", None)                              # :214
    qq, kk, vv = torch.split(qkv, [key_depth, key_depth, value_depth], dim=1)
    sa = {k[len("_query_self_attention."):]: v for k, v in p.items() if k.startswith("_query_self_attention.")}
    attn = O.query_self_attention(qq.reshape(N, heads, key_depth // heads, L), kk.reshape(N, heads, key_depth // heads, L),
                                  vv.reshape(N, heads, value_depth // heads, L), sa)                      # :215-221
    q = O.gelu(q + O.conv_bn_1d(attn, p, "_query_conv3_bn", "syncbn", None))                              # :222-224
    f = O.conv_bn_1d(O.conv_bn_1d(q, p, "_query_ffn_conv1_bn_act", "syncbn", "gelu"), p, "_query_ffn_conv2_bn", "syncbn", None)   # :227-228
    return O.gelu(q + f), pred                                                                            # :229-232
