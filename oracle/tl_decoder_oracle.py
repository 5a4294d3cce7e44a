"""TEST INFRASTRUCTURE (checker only; imported by tests/ and oracle scripts, never by the product).

CPU restatement of the Tube-Link mask decoder layer: `DetrTransformerDecoderLayer` with operation_order ('cross_attn', 'norm', 'self_attn',
'norm', 'ffn', 'norm') -- TL/mmdet/models/utils/transformer.py:408-451, configured in TL/configs/video/** (MultiheadAttention 256 / 8 heads,
batch_first=False; FFN 256 -> 2048 -> 256 ReLU with identity; post-norm) and called at
TL/models/video/tube_link_vis/mask2former_video_cc_head.py:883-894 with attn_masks = [attn_mask, None].

The classes it is built from (MultiheadAttention, BaseTransformerLayer, FFN) live in mmcv-full == 1.6.1, which is neither vendored in the
reference tree nor installed here: their wrapper semantics (query + query_pos, key + key_pos, value without position, identity residual,
boolean mask True = blocked, post-norm order) are restated from the published mmcv 1.x sources -- PARITY UNPINNED for the wrapper.  The
arithmetic core is pinned on torch.nn.MultiheadAttention / nn.LayerNorm / nn.Linear themselves (the modules mmcv wraps):
tests/test_oracle_golden.py::test_tl_decoder_layer_oracle_against_torch_mha.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch


def multihead_attention(query, key, value, identity, query_pos, key_pos, attn_mask, p: Dict[str, torch.Tensor], prefix: str, heads: int = 8):
    """identity + out_proj(softmax((q Wq + bq) / sqrt(d) (k Wk + bk)^T + mask) (v Wv + bv)); tensors are [N, B, E] (batch_first=False)."""
    E = query.shape[-1]
    d = E // heads
    w, b = p[prefix + "attn.in_proj_weight"], p[prefix + "attn.in_proj_bias"]
    q_in = query if query_pos is None else query + query_pos
    k_in = key if key_pos is None else key + key_pos
    q = q_in @ w[:E].T + b[:E]
    k = k_in @ w[E:2 * E].T + b[E:2 * E]
    v = value @ w[2 * E:].T + b[2 * E:]
    Nq, B, _ = q.shape
    L = k.shape[0]
    qh = q.reshape(Nq, B, heads, d).permute(1, 2, 0, 3) / math.sqrt(d)            # [B, h, Nq, d]
    kh = k.reshape(L, B, heads, d).permute(1, 2, 0, 3)
    vh = v.reshape(L, B, heads, d).permute(1, 2, 0, 3)
    s = qh @ kh.transpose(-1, -2)                                                  # [B, h, Nq, L]
    if attn_mask is not None:
        s = s.masked_fill(attn_mask.reshape(B, heads, Nq, L), float("-inf"))
    a = torch.softmax(s, dim=-1)
    o = (a @ vh).permute(2, 0, 1, 3).reshape(Nq, B, E)
    return identity + o @ p[prefix + "attn.out_proj.weight"].T + p[prefix + "attn.out_proj.bias"]


def _layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def decoder_layer(query, key, value, query_pos, key_pos, attn_masks: Optional[Sequence], p: Dict[str, torch.Tensor], heads: int = 8):
    masks = [None, None] if attn_masks is None else list(attn_masks)
    x = multihead_attention(query, key, value, query, query_pos, key_pos, masks[0], p, "attentions.0.", heads)
    x = _layer_norm(x, p["norms.0.weight"], p["norms.0.bias"])
    x = multihead_attention(x, x, x, x, query_pos, query_pos, masks[1], p, "attentions.1.", heads)
    x = _layer_norm(x, p["norms.1.weight"], p["norms.1.bias"])
    h = torch.relu(x @ p["ffns.0.layers.0.0.weight"].T + p["ffns.0.layers.0.0.bias"])
    x = x + h @ p["ffns.0.layers.1.weight"].T + p["ffns.0.layers.1.bias"]
    return _layer_norm(x, p["norms.2.weight"], p["norms.2.bias"])


def torch_module_composition(query, key, value, query_pos, key_pos, attn_masks, p: Dict[str, torch.Tensor], heads: int = 8):
    """The same layer built from the torch modules mmcv wraps (nn.MultiheadAttention, nn.LayerNorm, nn.Linear): the pin of the arithmetic."""
    import torch.nn as nn
    E = query.shape[-1]
    masks = [None, None] if attn_masks is None else list(attn_masks)

    def mha(prefix):
        m = nn.MultiheadAttention(E, heads, dropout=0.0).eval()
        m.load_state_dict({"in_proj_weight": p[prefix + "attn.in_proj_weight"], "in_proj_bias": p[prefix + "attn.in_proj_bias"],
                           "out_proj.weight": p[prefix + "attn.out_proj.weight"], "out_proj.bias": p[prefix + "attn.out_proj.bias"]})
        return m

    def ln(i):
        m = nn.LayerNorm(E).eval()
        m.load_state_dict({"weight": p[f"norms.{i}.weight"], "bias": p[f"norms.{i}.bias"]})
        return m

    with torch.no_grad():
        x = query + mha("attentions.0.")(query + query_pos, key + key_pos, value, attn_mask=masks[0], need_weights=False)[0]
        x = ln(0)(x)
        x = x + mha("attentions.1.")(x + query_pos, x + query_pos, x, attn_mask=masks[1], need_weights=False)[0]
        x = ln(1)(x)
        h = torch.relu(torch.nn.functional.linear(x, p["ffns.0.layers.0.0.weight"], p["ffns.0.layers.0.0.bias"]))
        x = x + torch.nn.functional.linear(h, p["ffns.0.layers.1.weight"], p["ffns.0.layers.1.bias"])
        return ln(2)(x)
