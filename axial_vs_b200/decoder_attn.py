"""Clip-level kMaX decoder attention, query side.

Drop-in for ``AttentionOperation`` (``Vk/maxtron_deeplab/modeling/transformer_decoder/maxtron_transformer_decoder.py:49-71``,
same constructor, forward signature and state-dict keys) and for the k-means cluster update inside
``kMaXTransformerLayer.forward`` (same file, lines 196-208).  Inference only; both run as fp32 CUDA kernels through the
C ABI (``axvs_query_self_attn``, ``axvs_kmeans_update``).  The pixel-side convolutions of the layer stay stock PyTorch.
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
