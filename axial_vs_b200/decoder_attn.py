"""Clip-level kMaX decoder attention, query side.

Drop-in for ``AttentionOperation`` (``Vk/maxtron_deeplab/modeling/transformer_decoder/maxtron_transformer_decoder.py:49-71``,
same constructor, forward signature and state-dict keys) and for the k-means cluster update inside
``kMaXTransformerLayer.forward`` (same file, lines 196-208).  Inference only; both run as fp32 CUDA kernels through the
C ABI (``axvs_query_self_attn``, ``axvs_kmeans_update``).  The second half of this file holds the WHOLE layer: ``kMaXPredictor`` and
``kMaXTransformerLayer`` drop-ins (pixel side included) built from those two kernels, the library GEMMs and the query x pixel contraction.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops

BN_EPS = 1e-3          # get_norm('syncbn', ...) -> nn.SyncBatchNorm(eps=1e-3, momentum=0.01), kmax_pixel_decoder.py:36-37


def _fold_bn(bn: nn.BatchNorm1d) -> torch.Tensor:
    """Eval-mode BatchNorm as interleaved (scale, shift) pairs, fp32 [channels, 2]."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return torch.stack([scale, shift], dim=1).contiguous()


class AttentionOperation(nn.Module):
    """forward(query, key, value): query/key [N, heads, 16, L], value [N, heads, 32, L] -> [N, heads*32, L]."""

    def __init__(self, channels_v: int, num_heads: int):
        super().__init__()
        # plain BatchNorm modules: same parameter / buffer names as the reference's SyncBatchNorm, eval statistics only
        self._batch_norm_similarity = nn.BatchNorm2d(num_heads, eps=BN_EPS, momentum=0.01)
        self._batch_norm_retrieved_value = nn.BatchNorm1d(channels_v, eps=BN_EPS, momentum=0.01)
        self._folded = None

    def _affines(self, device):
        bns = (self._batch_norm_similarity, self._batch_norm_retrieved_value)
        key = tuple((t._version, t.data_ptr()) for bn in bns for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var)) + (str(device),)
        if self._folded is None or self._folded[0] != key:
            self._folded = (key, _fold_bn(bns[0]).to(device), _fold_bn(bns[1]).to(device))
        return self._folded[1], self._folded[2]

    def forward(self, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise RuntimeError("axial_vs_b200.AttentionOperation is inference-only: call .eval() (training keeps the stock module)")
        sim, val = self._affines(query.device)
        return ops.query_self_attn(query.float().contiguous(), key.float().contiguous(), value.float().contiguous(), sim, val)


def kmeans_cluster_update(mask_logits: torch.Tensor, pixel_value: torch.Tensor, advanced_kmax: bool = False,
                          return_assignment: bool = False):
    """``mask_logits`` [N, L, (TH), W] or [N, L, M]; ``pixel_value`` [N, 256, M] -> ``kmeans_update`` [N, 256, L].

    Equivalent of ``index = logits.max(1)[1]; one_hot = zeros.scatter_(1, index, 1); einsum('blm,bdm->bdl', one_hot, value)``
    (optionally divided by the clamped pixel count per cluster).
    """
    ml = mask_logits.flatten(2).float().contiguous()
    pv = pixel_value.flatten(2).float().contiguous()
    return ops.kmeans_update(ml, pv, advanced=advanced_kmax, return_assignment=return_assignment)


# ------------------------------------------------------------------------------------------------------------------
# The whole clip-level decoder layer (row A11, Video-kMaX half): kMaXPredictor (DEC:75-124) and kMaXTransformerLayer (DEC:127-232),
# DEC = Vk/maxtron_deeplab/modeling/transformer_decoder/maxtron_transformer_decoder.py.  Same constructor arguments, forward signature,
# return values and state-dict keys (sub-module names `conv` / `norm` as in the reference's ConvBN, kmax_pixel_decoder.py:42-72).
# Every 1x1 ConvBN runs as ONE split-precision GEMM (fp32-grade products: the k-means step takes an argmax over the mask logits) over
# token rows with the eval-mode batch norm folded into its weights; the depthwise 5x5 ConvBN, the layout changes between the reference's
# channel-major tensors and token rows, the L2 normalisation and the residual + GELU steps are small kernels (csrc/kmax_layer.cuh); the
# query x pixel contraction, the k-means update and the query self-attention are the kernels of rows A10 / A11 above.  Inference only.
# ------------------------------------------------------------------------------------------------------------------
import math

from .modules import _PackedCache, _invalidate_hook, _require_inference


class ConvBN(nn.Module):
    """Parameter holder with the reference ConvBN's names (`conv.weight`, `conv.bias`, `norm.*`); eval statistics only."""

    def __init__(self, in_channels, out_channels, kernel_size=1, groups=1, padding=0, bias=True, norm=None, act=None, conv_type="2d", **_):
        super().__init__()
        conv = nn.Conv2d if conv_type == "2d" else nn.Conv1d
        self.conv = conv(in_channels, out_channels, kernel_size=kernel_size, padding=padding, groups=groups, bias=bias)
        bn = nn.BatchNorm2d if conv_type == "2d" else nn.BatchNorm1d
        self.norm = bn(out_channels, eps=BN_EPS, momentum=0.01) if norm == "syncbn" else nn.Identity()
        self.act_code = {None: 0, "relu": 1, "gelu": 2}[act]

    def folded(self, pad_to: int = 256):
        """(W', b') with the batch norm folded in: W' = W * scale, b' = bias * scale + shift; rows zero-padded to a multiple of `pad_to`."""
        w = self.conv.weight.detach().float().flatten(1)
        b = self.conv.bias.detach().float() if self.conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
        if isinstance(self.norm, nn.modules.batchnorm._BatchNorm):
            sc = _fold_bn(self.norm)
            w, b = w * sc[:, :1], b * sc[:, 0] + sc[:, 1]
        n = w.shape[0]
        n_pad = (n + pad_to - 1) // pad_to * pad_to
        if n_pad != n:
            w = torch.cat((w, torch.zeros(n_pad - n, w.shape[1], device=w.device)))
            b = torch.cat((b, torch.zeros(n_pad - n, device=b.device)))
        return w.contiguous(), b.contiguous()


def _pack_linear(cb: ConvBN, extra_bias: torch.Tensor = None):
    w, b = cb.folded()
    if extra_bias is not None:
        b = b.clone()
        b[:extra_bias.numel()] += extra_bias.to(b)
    return ops.pack_weight_split(w), b, w.shape[0], cb.act_code


def _lin(rows: torch.Tensor, packed) -> torch.Tensor:
    w, b, n_out, act = packed
    return ops.linear_f32(rows, w, b, n_out, act=act, split=True)


class kMaXPredictor(nn.Module):
    """DEC:75-124.  forward(mask_embeddings [N, C, L], class_embeddings [N, C, L], pixel_feature [N, C, TH, W]) -> dict."""

    def __init__(self, in_channel_pixel, in_channel_query, num_classes=133 + 1):
        super().__init__()
        self._pixel_space_head_conv0bnact = ConvBN(in_channel_pixel, in_channel_pixel, kernel_size=5, groups=in_channel_pixel, padding=2, bias=False,
                                                   norm="syncbn", act="gelu")
        self._pixel_space_head_conv1bnact = ConvBN(in_channel_pixel, 256, kernel_size=1, bias=False, norm="syncbn", act="gelu")
        self._pixel_space_head_last_convbn = ConvBN(256, 128, kernel_size=1, bias=True, norm="syncbn", act=None)
        self._transformer_mask_head = ConvBN(256, 128, kernel_size=1, bias=False, norm="syncbn", act=None, conv_type="1d")
        self._transformer_class_head = ConvBN(256, num_classes, kernel_size=1, norm=None, act=None, conv_type="1d")
        self._pixel_space_mask_batch_norm = nn.BatchNorm2d(1, eps=BN_EPS, momentum=0.01)
        self._num_classes = num_classes
        self._cache = _PackedCache()
        self.register_load_state_dict_post_hook(_invalidate_hook)

    def _packed(self, dev):
        def build():
            K = self._num_classes
            void = torch.zeros(K)
            void[-1] = math.log((K - 1) * 0.9 / (1 - 0.9))                # add_bias_towards_void (DEC:37-45), a constant: folded into the bias
            dw = self._pixel_space_head_conv0bnact
            bn1 = self._pixel_space_mask_batch_norm
            sc = float(bn1.weight.detach() / torch.sqrt(bn1.running_var.detach() + bn1.eps))
            return {"dw_w": dw.conv.weight.detach().float().reshape(-1, 25).contiguous(), "dw_affine": _fold_bn(dw.norm),
                    "c1": _pack_linear(self._pixel_space_head_conv1bnact), "last": _pack_linear(self._pixel_space_head_last_convbn),
                    "mask": _pack_linear(self._transformer_mask_head), "cls": _pack_linear(self._transformer_class_head, void),
                    "bn_scale": sc, "bn_shift": float(bn1.bias.detach() - bn1.running_mean.detach() * sc)}
        return self._cache.get(self, dev, build)

    def forward_rows(self, query_rows: torch.Tensor, pixel_rows: torch.Tensor, N: int, L: int, TH: int, W: int):
        """query_rows fp32 [N*L, C], pixel_rows fp32 [N*TH*W, C] (channels-last token rows)."""
        pk = self._packed(pixel_rows.device)
        M = TH * W
        f = ops.dwconv5(pixel_rows, pk["dw_w"], pk["dw_affine"], N, TH, W, act=2)
        f = _lin(f, pk["c1"])
        f = _lin(f, pk["last"])                                           # [N*M, 256], the first 128 columns are the embedding
        pixel_norm = ops.rows_to_cm(f, N, 128, normalize=True)            # [N, 128, M]
        cls = _lin(query_rows, pk["cls"]).view(N, L, -1)[:, :, :self._num_classes]
        mk = _lin(query_rows, pk["mask"])                                 # [N*L, 256]: mask kernels, row (n, l), first 128 columns
        mask_logits = torch.empty(N, L, M, dtype=torch.float32, device=pixel_rows.device)
        for n in range(N):                                                # [L, 1, M] per clip lands in place (the kernel's output is query-major)
            ops.mask_einsum(pixel_norm[n], mk[n * L:(n + 1) * L], 1, L, M, pk["bn_scale"], pk["bn_shift"], out=mask_logits[n])
        return {"class_logits": cls.contiguous(), "mask_logits": mask_logits.view(N, L, TH, W),
                "mask_embeddings": mk.view(N, L, -1)[:, :, :128].contiguous(), "pixel_feature": pixel_norm.view(N, 128, TH, W)}

    def forward(self, mask_embeddings, class_embeddings, pixel_feature):
        _require_inference(self, pixel_feature)
        if mask_embeddings is not class_embeddings and not torch.equal(mask_embeddings, class_embeddings):
            raise NotImplementedError("kMaXPredictor: the layer passes the same tensor as mask and class embeddings (DEC:195-196)")
        N, C, TH, W = pixel_feature.shape
        L = mask_embeddings.shape[2]
        return self.forward_rows(ops.cm_to_rows(mask_embeddings.float().contiguous()), ops.cm_to_rows(pixel_feature.float().reshape(N, C, TH * W).contiguous()),
                                 N, L, TH, W)


class kMaXTransformerLayer(nn.Module):
    """DEC:127-232.  forward(pixel_feature [N, C, TH, W], query_feature [N, 256, L]) -> (query_feature, prediction_result)."""

    def __init__(self, num_classes=133, in_channel_pixel=2048, in_channel_query=256, base_filters=128, num_heads=8, bottleneck_expansion=2,
                 key_expansion=1, value_expansion=2, drop_path_prob=0.0, advanced_kmax=False, skip_conn_init_value=0.0):
        super().__init__()
        self._num_classes, self._num_heads = num_classes, num_heads
        self._bottleneck_channels = int(round(base_filters * bottleneck_expansion))
        self._total_key_depth = int(round(base_filters * key_expansion))
        self._total_value_depth = int(round(base_filters * value_expansion))
        if (self._bottleneck_channels, self._total_key_depth, self._total_value_depth, in_channel_query, num_heads) != (256, 128, 256, 256, 8):
            raise NotImplementedError("kMaXTransformerLayer: built for the shipped configuration (bottleneck 256, key depth 128, value depth 256, 8 heads)")
        self.advanced_kmax = advanced_kmax
        B, Kd, Vd = self._bottleneck_channels, self._total_key_depth, self._total_value_depth
        self._query_conv1_bn_act = ConvBN(in_channel_query, B, bias=False, norm="syncbn", act="gelu", conv_type="1d")
        self._pixel_conv1_bn_act = ConvBN(in_channel_pixel, B, bias=False, norm="syncbn", act="gelu")
        self._query_qkv_conv_bn = ConvBN(B, Kd * 2 + Vd, bias=False, norm="syncbn", act=None, conv_type="1d")
        self._pixel_v_conv_bn = ConvBN(B, Vd, bias=False, norm="syncbn", act=None)
        self._query_self_attention = AttentionOperation(channels_v=Vd, num_heads=num_heads)
        self._query_conv3_bn = ConvBN(Vd, in_channel_query, bias=False, norm="syncbn", act=None, conv_type="1d")
        self._query_ffn_conv1_bn_act = ConvBN(in_channel_query, 2048, bias=False, norm="syncbn", act="gelu", conv_type="1d")
        self._query_ffn_conv2_bn = ConvBN(2048, in_channel_query, bias=False, norm="syncbn", act=None, conv_type="1d")
        self._predictor = kMaXPredictor(in_channel_pixel=B, in_channel_query=B, num_classes=num_classes)
        self._kmeans_query_batch_norm_retrieved_value = nn.BatchNorm1d(Vd, eps=BN_EPS, momentum=0.01)
        self._kmeans_query_conv3_bn = ConvBN(Vd, in_channel_query, bias=False, norm="syncbn", act=None, conv_type="1d")
        self._cache = _PackedCache()
        self.register_load_state_dict_post_hook(_invalidate_hook)

    def _packed(self, dev):
        def build():
            # the batch norm on the k-means update (DEC:209) is an affine per INPUT channel of the 1x1 conv that follows (:210): fold it there
            km = self._kmeans_query_conv3_bn
            w, b = km.folded()
            sc = _fold_bn(self._kmeans_query_batch_norm_retrieved_value)
            b = b + w @ sc[:, 1]
            w = (w * sc[:, 0][None, :]).contiguous()
            return {"q1": _pack_linear(self._query_conv1_bn_act), "p1": _pack_linear(self._pixel_conv1_bn_act), "qkv": _pack_linear(self._query_qkv_conv_bn),
                    "pv": _pack_linear(self._pixel_v_conv_bn), "q3": _pack_linear(self._query_conv3_bn), "f1": _pack_linear(self._query_ffn_conv1_bn_act),
                    "f2": _pack_linear(self._query_ffn_conv2_bn), "km": (ops.pack_weight_split(w), b.contiguous(), w.shape[0], 0)}
        return self._cache.get(self, dev, build)

    def forward(self, pixel_feature: torch.Tensor, query_feature: torch.Tensor):
        _require_inference(self, pixel_feature, query_feature)
        N, C, TH, W = pixel_feature.shape
        _, D, L = query_feature.shape
        M = TH * W
        pk = self._packed(pixel_feature.device)
        Kd, Vd, H = self._total_key_depth, self._total_value_depth, self._num_heads
        # pixel / query spaces (DEC:186-187): GELU on the pixel feature folded into the layout change
        pixel_space = _lin(ops.cm_to_rows(pixel_feature.float().reshape(N, C, M).contiguous(), gelu=True), pk["p1"])      # [N*M, 256]
        q_rows0 = ops.cm_to_rows(query_feature.float().contiguous())                                                       # [N*L, 256]
        query_space = _lin(q_rows0, pk["q1"])
        # k-means cross-attention (DEC:190-211)
        pixel_value = ops.rows_to_cm(_lin(pixel_space, pk["pv"]), N, Vd)                                                   # [N, 256, M]
        pred = self._predictor.forward_rows(query_space, pixel_space, N, L, TH, W)
        km = ops.kmeans_update(pred["mask_logits"].view(N, L, M), pixel_value, advanced=self.advanced_kmax)                # [N, 256, L]
        q_rows = ops.add_act(q_rows0, _lin(ops.cm_to_rows(km), pk["km"]), act=0)
        # query self-attention (DEC:214-225)
        qkv = _lin(query_space, pk["qkv"])                                                                                 # [N*L, 512]
        qkv_cm = torch.cat([ops.rows_to_cm(qkv[:, i * 256:(i + 1) * 256].contiguous(), N, 256) for i in range(2)], dim=1)  # [N, 512, L]
        q, k, v = torch.split(qkv_cm, [Kd, Kd, Vd], dim=1)
        attn = self._query_self_attention(q.reshape(N, H, Kd // H, L), k.reshape(N, H, Kd // H, L), v.reshape(N, H, Vd // H, L))
        q_rows = ops.add_act(q_rows, _lin(ops.cm_to_rows(attn), pk["q3"]), act=2)
        # FFN (DEC:228-231)
        q_rows = ops.add_act(q_rows, _lin(_lin(q_rows, pk["f1"]), pk["f2"]), act=2)
        return ops.rows_to_cm(q_rows, N, D), pred
