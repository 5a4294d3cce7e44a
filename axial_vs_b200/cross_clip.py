"""Drop-in modules for the cross-clip tracking module's trajectory attention (Video-kMaX flavour).

Mirrors `MaXTron_Video-kMaX/maxtron_deeplab/modeling/cross_clip_tracking_module/maxtron_cross_clip_tracking_module.py`
(`CC`): TrajectoryAttention :78-130 (fused `qkv` Linear, no positional term) and TrajectoryAttentionLayer :133-173
(post-LN residual block), with identical constructor arguments, forward signatures and state-dict keys
(`self_attn.{qkv,proj_q,proj_kv,proj}.{weight,bias}`, `norm.{weight,bias}`).  The Tube-Link copy
(`TL/models/video/tube_link_vis/mask2former_video_cc_head.py:152-247`) has the same math and leaf names.

Frames of the trajectory attention are CLIPS here (num_frames = T_clips) and the tokens of a frame are the
`seq_len` object queries of that clip.  Inference only; no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .modules import _PackedCache, _require_inference


class TrajectoryAttention(nn.Module):
    """CC:78-130.  forward(x [b, T*Q, C], seq_len=Q, num_frames=T) -> x."""

    def __init__(self, d_model, nhead, attn_drop):
        super().__init__()
        if d_model != ops.C or nhead != ops.HEADS:
            raise NotImplementedError(f"axial_vs_b200 kernels are specialised for d_model=256, nhead=8 (got {d_model}, {nhead})")
        self.num_heads = nhead
        self.head_dim = d_model // nhead
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(d_model, d_model * 3)
        self.proj_q = nn.Linear(d_model, d_model)
        self.proj_kv = nn.Linear(d_model, d_model * 2)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(d_model, d_model)
        self._cache = _PackedCache()

    def packed(self, device) -> ops.PackedTA:
        return self._cache.get(self, device, lambda: ops.pack_ta(dict(self.state_dict())))

    def _run(self, x: Tensor, seq_len: int, num_frames: int, residual: bool) -> Tensor:
        _require_inference(self, x)
        b, N, C = x.shape
        if N != seq_len * num_frames:
            raise RuntimeError(f"sequence length {N} != seq_len {seq_len} * num_frames {num_frames}")
        xf = x.contiguous().float().view(b * N, C)
        out = ops.traj_attn_fwd(xf, xf, xf, None, xf if residual else None, self.packed(x.device), b, num_frames, seq_len, 1, ops.AXIS_NONE)
        return out.view(b, N, C)

    def forward(self, x, seq_len=128, num_frames=6):
        return self._run(x, seq_len, num_frames, residual=False).to(x.dtype)


class TrajectoryAttentionLayer(nn.Module):
    """CC:133-173.  forward(x, seq_len, num_frames) = LayerNorm(x + self_attn(x))  (normalize_before=False everywhere)."""

    def __init__(self, d_model, nhead, dropout=0.0, attn_drop=0.0, activation="relu", normalize_before=False):
        super().__init__()
        self.self_attn = TrajectoryAttention(d_model, nhead, attn_drop=attn_drop)
        self.norm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.normalize_before = normalize_before
        self._reset_parameters()

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward_post(self, tgt, seq_len, num_frames):
        s = self.self_attn._run(tgt, seq_len, num_frames, residual=True)          # tgt + attn(tgt), fused residual epilogue
        out = ops.layernorm(s.view(-1, ops.C), self.norm.weight.detach().float(), self.norm.bias.detach().float(), self.norm.eps)
        return out.view_as(s).to(tgt.dtype)

    def forward_pre(self, tgt, seq_len, num_frames):
        # reference quirk (CC:163-168): the pre-norm result is computed and discarded -> plain residual attention
        return self.self_attn._run(tgt, seq_len, num_frames, residual=True).to(tgt.dtype)

    def forward(self, x, seq_len, num_frames):
        if self.normalize_before:
            return self.forward_pre(x, seq_len, num_frames)
        return self.forward_post(x, seq_len, num_frames)
