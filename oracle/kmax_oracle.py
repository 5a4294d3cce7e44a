"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the kMaX pixel decoder's axial attention.

Follows `AxialAttention` / `AxialAttention2D`, Vk/kmax_deeplab/modeling/pixel_decoder/kmax_pixel_decoder.py:76-190 (eval mode: the
SyncBatchNorm layers apply their running statistics).  Pinned against the reference itself: `oracle/make_golden.py kmax` runs the
unmodified modules (imported through `oracle/ref_loader.cross_clip`, which loads that file) and stores their outputs in
`tests/golden/kmax_axial_*.npz`; `tests/test_oracle_golden.py` replays them through this file.  Only tests/, `__graft_entry__.smoke()`
and bench.py's CPU legs may import it.
"""
from __future__ import annotations

import torch

MAX_SPAN = 255          # :76
BN_EPS = 1e-3           # get_norm('syncbn'): nn.SyncBatchNorm(channels, eps=1e-3), :37


def bn_affine(p: dict, prefix: str):
    """Eval-mode batch norm as y = x * s + t."""
    s = p[prefix + ".weight"].double() / torch.sqrt(p[prefix + ".running_var"].double() + BN_EPS)
    return s, p[prefix + ".bias"].double() - p[prefix + ".running_mean"].double() * s


def rpe_table(emb: torch.Tensor, L: int) -> torch.Tensor:
    """RelativePositionalEncoding.forward (:77-102) for query_length = key_length = L: table[l, m] = emb[m - l + MAX_SPAN - 1]."""
    idx = torch.arange(L)[None, :] - torch.arange(L)[:, None] + MAX_SPAN - 1
    return emb[idx]                                                  # [L, L, depth]


def axial_attention(x: torch.Tensor, p: dict, num_heads: int = 8) -> torch.Tensor:
    """AxialAttention.forward (:128-157).  x [N, C, L] -> [N, total_value_depth, L]; computed in float64 per (sequence, head)."""
    N, C, L = x.shape
    w = p["qkv_transform.conv.weight"].double()[:, :, 0]             # [2 Kd + Vd, C]
    emb_q, emb_k, emb_v = (p[f"_{n}_rpe._embeddings.weight"].double() for n in ("query", "key", "value"))
    dk, dv = emb_q.shape[1], emb_v.shape[1]
    Kd, Vd = dk * num_heads, dv * num_heads
    s_qkv, t_qkv = bn_affine(p, "_batch_norm_qkv")
    s_sim, t_sim = bn_affine(p, "_batch_norm_similarity")            # channels: [content h0..h7, query-rpe h0..h7, key-rpe h0..h7]
    s_out, t_out = bn_affine(p, "_batch_norm_retrieved_output")      # channels: [content (h, d), rpe (h, d)]
    tok = x.double().permute(0, 2, 1)                                # [N, L, C] tokens
    qkv = tok @ w.t() * s_qkv + t_qkv                                # [N, L, 2 Kd + Vd]
    rq, rk, rv = rpe_table(emb_q, L), rpe_table(emb_k, L), rpe_table(emb_v, L)
    out = torch.empty(N, Vd, L, dtype=torch.float64)
    H = num_heads
    for h in range(H):
        q = qkv[:, :, h * dk:(h + 1) * dk]                           # [N, L, dk]
        k = qkv[:, :, Kd + h * dk:Kd + (h + 1) * dk]
        v = qkv[:, :, 2 * Kd + h * dv:2 * Kd + (h + 1) * dv]         # [N, L, dv]
        content = q @ k.transpose(1, 2)                              # [N, L(l), L(m)]
        q_rpe = (q[:, :, None, :] * rq[None]).sum(-1)                # q_l . rq[l, m]
        k_rpe = (k[:, None, :, :] * rk[None]).sum(-1)                # k_m . rk[l, m]
        logits = (content * s_sim[h] + t_sim[h]) + (q_rpe * s_sim[H + h] + t_sim[H + h]) + (k_rpe * s_sim[2 * H + h] + t_sim[2 * H + h])
        wts = torch.softmax(logits, dim=-1)
        got_c = wts @ v                                              # [N, L, dv]
        got_r = (wts[:, :, :, None] * rv[None]).sum(2)               # sum_m w[l, m] rv[l, m, :]
        cs = slice(h * dv, (h + 1) * dv)
        rs = slice(Vd + h * dv, Vd + (h + 1) * dv)
        y = got_c * s_out[cs] + t_out[cs] + got_r * s_out[rs] + t_out[rs]
        out[:, cs, :] = y.transpose(1, 2)
    return out.float()


def axial_attention_2d(x: torch.Tensor, p_height: dict, p_width: dict, num_heads: int = 8) -> torch.Tensor:
    """AxialAttention2D.forward (:177-190): height axis over (N W) sequences, then width axis over (N H) sequences."""
    N, C, Hh, Ww = x.shape
    y = axial_attention(x.permute(0, 3, 1, 2).reshape(N * Ww, C, Hh), p_height, num_heads)            # [(N W), Vd, H]
    Vd = y.shape[1]
    y = y.reshape(N, Ww, Vd, Hh).permute(0, 3, 2, 1).reshape(N * Hh, Vd, Ww)
    y = axial_attention(y, p_width, num_heads)                                                        # [(N H), Vd, W]
    return y.reshape(N, Hh, Vd, Ww).permute(0, 2, 1, 3).contiguous()
