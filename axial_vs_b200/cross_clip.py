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
from .modules import _PackedCache, _invalidate_hook, _require_inference


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
        self._cache_split = _PackedCache()
        self.precision = "bf16"            # "bf16": the fused bf16 kernels (default); "split": fp32-grade products, see _run_split
        self.register_load_state_dict_post_hook(_invalidate_hook)

    def packed(self, device) -> ops.PackedTA:
        return self._cache.get(self, device, lambda: ops.pack_ta(dict(self.state_dict())))

    def _packed_split(self, device):
        def build():
            C = ops.C
            wqkv, bqkv = self.qkv.weight.detach().float(), self.qkv.bias.detach().float()
            wkv, bkv = self.proj_kv.weight.detach().float(), self.proj_kv.bias.detach().float()
            lin = lambda w, b: (ops.pack_weight_split(w.contiguous()), b.contiguous())
            return {"q": lin(wqkv[:C], bqkv[:C]), "k": lin(wqkv[C:2 * C], bqkv[C:2 * C]), "v": lin(wqkv[2 * C:], bqkv[2 * C:]),
                    "pq": lin(self.proj_q.weight.detach().float(), self.proj_q.bias.detach().float()),
                    "k2": lin(wkv[:C], bkv[:C]), "v2": lin(wkv[C:], bkv[C:]),
                    "proj": lin(self.proj.weight.detach().float(), self.proj.bias.detach().float())}
        return self._cache_split.get(self, device, build)

    def _run_split(self, x: Tensor, seq_len: int, num_frames: int, residual: bool) -> Tensor:
        """The same math (CC:95-130) at fp32-grade accuracy: every Linear as a split-precision GEMM (bf16 hi / lo operands, three tensor-core
        products: axvs_linear_f32), both attentions in fp32 on the masked-attention kernel (per-frame softmax = key splits on the frames;
        attention over the frames = one query with F keys per token).  About 5x the time of the fused bf16 kernels on this small stage;
        it exists because the final per-pixel argmax over 128 nearly tied query logits resolves below the bf16-compute error
        (99.85 % label agreement with bf16 attention, >= 99.9 % with this path)."""
        b, N, C = x.shape
        F_, n = num_frames, seq_len
        pk = self._packed_split(x.device)
        rows = x.contiguous().float().view(b * N, C)
        lin = lambda a, key, scale=1.0: ops.linear_f32(a, pk[key][0], pk[key][1], C, act=0, split=True, scale=scale)
        s2 = self.scale * 1.4426950408889634                   # softmax scale and log2(e): the attention kernel works in the exp2 domain
        q, k, v = lin(rows, "q", s2).view(b, N, C), lin(rows, "k").view(b, N, C), lin(rows, "v").view(b, N, C)
        xs = ops.frame_attn_f32(q, k, v, F_, heads=self.num_heads)                    # [b, N, F, C]
        own = (torch.arange(N, device=x.device) // n).view(1, N, 1, 1).expand(b, N, 1, C)
        x_diag = torch.gather(xs, 2, own).view(b * N, C)                              # the token's own frame (CC:112-114); pure indexing
        q2 = lin(x_diag, "pq", s2)                                                    # (proj_q(x_diag)) * scale, CC:115-120
        xf = xs.view(b * N * F_, C)
        k2, v2 = lin(xf, "k2"), lin(xf, "v2")
        o = torch.empty(b * N, C, dtype=torch.float32, device=x.device)
        for r0 in range(0, b * N, 32768):                                             # one "batch element" per token: grid limits
            r1 = min(b * N, r0 + 32768)
            o[r0:r1] = ops.masked_mha(q2[r0:r1].view(r1 - r0, 1, C), k2[r0 * F_:r1 * F_].view(r1 - r0, F_, C), v2[r0 * F_:r1 * F_].view(r1 - r0, F_, C),
                                      None, heads=self.num_heads, seq_first=False, out_dtype=torch.float32).view(r1 - r0, C)
        out = lin(o, "proj")
        return (ops.add_act(out, rows, act=0) if residual else out).view(b, N, C)

    def _run(self, x: Tensor, seq_len: int, num_frames: int, residual: bool) -> Tensor:
        _require_inference(self, x)
        b, N, C = x.shape
        if N != seq_len * num_frames:
            raise RuntimeError(f"sequence length {N} != seq_len {seq_len} * num_frames {num_frames}")
        if self.precision == "split":
            return self._run_split(x, seq_len, num_frames, residual)
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


# ------------------------------------------------------------------------------------------------------------------
# The rest of the cross-clip tracking module (CC:30-75, 176-331): temporal ASPP, embedding projections, predictor.
# Parameter names mirror the reference state dict (`conv_short_aggregate_layers.{i}._aspp_conv{j}`, `conv_norms.{i}`,
# `_class_embedding_projection.{conv,norm}`, `_mask_embedding_projection.{conv,norm}`, `_predictor.*`).
# ------------------------------------------------------------------------------------------------------------------
import math
from typing import List


class _ConvBN1d(nn.Module):
    """`ConvBN(conv_type='1d', kernel_size=1)` (Vk/kmax_deeplab/modeling/pixel_decoder/kmax_pixel_decoder.py:42-72), eval only.
    The 1x1 conv + eval-mode (Sync)BatchNorm is folded into one GEMM: W' = W * g/sqrt(rv+eps), b' = beta - rm*g/sqrt(rv+eps) (+ conv bias)."""

    def __init__(self, cin, cout, bias=True, norm=None, act=None):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, kernel_size=1, bias=bias)
        self.norm = nn.BatchNorm1d(cout, eps=1e-3, momentum=0.01) if norm == "syncbn" else nn.Identity()
        self.act_code = {None: 0, "relu": 1, "gelu": 2}[act]
        self.cout = cout
        self._cache = _PackedCache()
        self.register_load_state_dict_post_hook(_invalidate_hook)
        self._cache32 = _PackedCache()

    def folded(self):
        w = self.conv.weight.detach().float()[:, :, 0]
        b = self.conv.bias.detach().float() if self.conv.bias is not None else torch.zeros(self.cout, device=w.device)
        if isinstance(self.norm, nn.BatchNorm1d):
            sc = self.norm.weight.detach().float() / torch.sqrt(self.norm.running_var.float() + self.norm.eps)
            w = w * sc[:, None]
            b = (b - self.norm.running_mean.float()) * sc + self.norm.bias.detach().float()
        return w, b

    def packed(self, device):
        def build():
            w, b = self.folded()
            n_pad = (self.cout + 255) // 256 * 256                      # the GEMM works on 256-column chunks
            wp = torch.zeros(n_pad, w.shape[1], device=w.device)
            wp[: self.cout] = w
            bp = torch.zeros(n_pad, device=w.device)
            bp[: self.cout] = b
            return ops.pack_weight(wp.contiguous()), bp.contiguous(), n_pad
        key_mod = nn.ModuleList([self.conv, self.norm]) if isinstance(self.norm, nn.BatchNorm1d) else self.conv
        return self._cache.get(key_mod, device, build)

    def packed_split(self, device):
        """Split-precision image of the folded weight (see ops.pack_weight_split) for `run32`."""
        def build():
            w, b = self.folded()
            n_pad = (self.cout + 255) // 256 * 256
            wp = torch.zeros(n_pad, w.shape[1], device=w.device)
            wp[: self.cout] = w
            bp = torch.zeros(n_pad, device=w.device)
            bp[: self.cout] = b
            return ops.pack_weight_split(wp), bp.contiguous(), n_pad
        key_mod = nn.ModuleList([self.conv, self.norm]) if isinstance(self.norm, nn.BatchNorm1d) else self.conv
        return self._cache32.get(key_mod, device, build)

    def run32(self, a_f32: Tensor, out_dtype=torch.float32) -> Tensor:
        """fp32 rows in, fp32-grade GEMM (split precision), fp32 (or bf16) rows out."""
        wp, bp, n_pad = self.packed_split(a_f32.device)
        return ops.linear_f32(a_f32, wp, bp, n_pad, self.act_code, True, out_dtype)

    def run(self, a_bf16: Tensor, out_dtype=torch.bfloat16, extra_bias: Tensor = None) -> Tensor:
        wp, bp, n_pad = self.packed(a_bf16.device)
        if extra_bias is not None:
            bp = bp.clone()
            bp[: self.cout] += extra_bias
        return ops.linear_act(a_bf16, wp, bp, n_pad, self.act_code, out_dtype)


class ASPP(nn.Module):
    """CC:176-201 (parameters only; the arithmetic is fused with the residual + conv_norms LayerNorm in `axvs_cc_aspp_fwd`)."""

    def __init__(self, in_channels, output_channels, kernel_sizes, atrous_rates, dropout_rate, norm_fn):
        super().__init__()
        if norm_fn != "ln" or list(kernel_sizes) != [3, 3, 3] or in_channels != 256 or output_channels != 256:
            raise NotImplementedError("axial_vs_b200: ASPP supports 256 channels, kernel_sizes [3,3,3], norm_fn 'ln' (every shipped config)")
        for j in range(3):
            setattr(self, f"_aspp_conv{j}", nn.Conv1d(in_channels, output_channels, kernel_size=3, dilation=atrous_rates[j], padding="same",
                                                      padding_mode="replicate"))
        self._proj_conv_bn_act = nn.Module()
        self._proj_conv_bn_act.conv = nn.Conv1d(output_channels * 3, output_channels, kernel_size=1, bias=False)
        self._proj_conv_bn_act.norm = nn.Module()
        self._proj_conv_bn_act.norm.weight = nn.Parameter(torch.ones(output_channels))
        self._proj_conv_bn_act.norm.bias = nn.Parameter(torch.zeros(output_channels))
        self.atrous_rates = list(atrous_rates)


class MaXTronCCPredictor(nn.Module):
    """CC:30-75, eval branch."""

    def __init__(self, num_classes=133 + 1):
        super().__init__()
        self._transformer_mask_head = _ConvBN1d(256, 128, bias=False, norm="syncbn")
        self._transformer_class_head = _ConvBN1d(256, num_classes)
        self._transformer_class_activation_head = _ConvBN1d(256, 1)
        self._pixel_space_mask_batch_norm = nn.BatchNorm1d(1, eps=1e-3, momentum=0.01)
        nn.init.constant_(self._pixel_space_mask_batch_norm.weight, 0.1)
        self.num_classes = num_classes
        self._scalars = _PackedCache()

    def _host_scalars(self, device):
        """(class-activation bias, folded pixel-space BN scale, shift) as Python floats: ONE host read per weight update, not one per call."""
        def build():
            act, bn = self._transformer_class_activation_head, self._pixel_space_mask_batch_norm
            sc = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
            sh = bn.bias.detach().float() - bn.running_mean.float() * sc
            v = torch.stack((act.conv.bias.detach().float().reshape(()), sc.reshape(()), sh.reshape(()))).tolist()
            return tuple(float(x) for x in v)
        return self._scalars.get(nn.ModuleList([self._transformer_class_activation_head.conv, self._pixel_space_mask_batch_norm]), device, build)

    def class_logits(self, class_embeddings, num_clips):
        """CC:47-51: class activation softmax over the clips, pooled class embedding, class head + void bias -> [1, Q, K+1]."""
        T = num_clips
        Q = class_embeddings.shape[0] // T
        act = self._transformer_class_activation_head
        pooled = ops.cc_class_pool(class_embeddings, act.conv.weight.detach().float().reshape(-1).contiguous(),
                                   self._host_scalars(class_embeddings.device)[0], T, Q)                     # [Q, 256]
        void = torch.zeros(self.num_classes, device=pooled.device)
        void[-1] = math.log((self.num_classes - 1) * 0.9 / (1 - 0.9))                                         # add_bias_towards_void
        cls = self._transformer_class_head.run(pooled, torch.float32, extra_bias=void)[:, : self.num_classes]  # [Q, K+1]
        return cls.unsqueeze(0)

    def mask_kernels(self, mask_embeddings):
        """CC:53: [T*Q, 256] rows (t, q), 128 valid columns; fp32 embeddings -> fp32 kernels through the split-precision GEMM."""
        if mask_embeddings.dtype == torch.float32:
            return self._transformer_mask_head.run32(mask_embeddings)
        return self._transformer_mask_head.run(mask_embeddings)

    def mask_logits(self, mk, pixel_feature, num_clips):
        """CC:62-68 for `num_clips` clips: mk rows (t, q) of those clips, pixel_feature fp32 [T, 128, V*H, W] -> fp32 [Q, T, V*H*W]."""
        T = num_clips
        Q = mk.shape[0] // T
        _, sc, sh = self._host_scalars(mk.device)
        _, Cp, VH, Wd = pixel_feature.shape
        return ops.mask_einsum(pixel_feature.contiguous().float(), mk.contiguous(), T, Q, VH * Wd, sc, sh)

    def forward(self, mask_embeddings, class_embeddings, pixel_feature, num_clips, num_clip_frames):
        """mask embeddings fp32 (or bf16), class embeddings bf16: [T*Q, 256] rows (t, q); pixel_feature fp32 [T, 128, V*H, W]."""
        T = num_clips
        Q = class_embeddings.shape[0] // T
        cls = self.class_logits(class_embeddings, T)
        ml = self.mask_logits(self.mask_kernels(mask_embeddings), pixel_feature, T)                            # [Q, T, P]
        _, Cp, VH, Wd = pixel_feature.shape
        V = num_clip_frames
        return {"class_logits": cls, "mask_logits": ml.view(1, Q, T * V, VH // V, Wd)}


class CrossClipTrackingModule(nn.Module):
    """CC:204-331.  forward(clip_query [1, Q, T, C], panoptic_features [1, 128, T*V, H, W]) -> dict (eval, batch of one video)."""

    def __init__(self, *, num_layers: int, num_classes: int, attn_drop: float, aspp_drop: float, kernel_sizes: List[int],
                 atrous_rates: List[int], norm_fn: str, num_clip_frames: int):
        super().__init__()
        self.kernel_sizes, self.atrous_rates = kernel_sizes, atrous_rates
        self.attn_drop, self.aspp_drop, self.norm_fn, self.num_clip_frames = attn_drop, aspp_drop, norm_fn, num_clip_frames
        self.num_heads = 8
        self.num_layers = num_layers
        self.transformer_trajectory_self_attention_layers = nn.ModuleList()
        self.conv_short_aggregate_layers = nn.ModuleList()
        self.conv_norms = nn.ModuleList()
        for _ in range(num_layers):
            self.transformer_trajectory_self_attention_layers.append(
                TrajectoryAttentionLayer(d_model=256, nhead=8, dropout=0.0, attn_drop=attn_drop, normalize_before=False))
            self.conv_short_aggregate_layers.append(ASPP(256, 256, kernel_sizes, atrous_rates, aspp_drop, norm_fn))
            self.conv_norms.append(nn.LayerNorm(256))
        self._class_embedding_projection = _ConvBN1d(256, 256, bias=False, norm="syncbn", act="gelu")
        self._mask_embedding_projection = _ConvBN1d(256, 256, bias=False, norm="syncbn", act="gelu")
        self._predictor = MaXTronCCPredictor(num_classes=num_classes + 1)
        self._aspp_cache = [_PackedCache() for _ in range(num_layers)]
        # The reference's eval branch moves class / mask logits to the host (CC:57,71: `.cpu()`), because its post-processing runs there.
        # This drop-in keeps them on the GPU by default (postprocess.PanopticPostProcessor consumes them in place); set
        # `outputs_on_cpu = True` for callers that mix the outputs with host tensors.
        self.outputs_on_cpu = False

    def set_precision(self, mode: str) -> "CrossClipTrackingModule":
        """"bf16" (default): the trajectory attention of the cross-clip layers runs on the fused bf16 kernels; "split": at fp32-grade accuracy
        (TrajectoryAttention._run_split).  Everything after the attention (temporal ASPP, projections, query x pixel contraction) is
        split-precision in both modes."""
        if mode not in ("bf16", "split"):
            raise ValueError("precision must be 'bf16' or 'split'")
        for layer in self.transformer_trajectory_self_attention_layers:
            layer.self_attn.precision = mode
        return self

    def _packed_aspp(self, i, device):
        mods = nn.ModuleList([self.conv_short_aggregate_layers[i], self.conv_norms[i]])
        return self._aspp_cache[i].get(mods, device, lambda: ops.pack_aspp(dict(self.conv_short_aggregate_layers[i].state_dict()),
                                                                            self.conv_norms[i].weight, self.conv_norms[i].bias, self.atrous_rates))

    def forward(self, clip_query, panoptic_features):
        _require_inference(self, clip_query, panoptic_features)
        b, Q, T, C = clip_query.shape
        if b != 1:
            raise NotImplementedError("axial_vs_b200: the cross-clip module runs one video at a time at inference (as the reference does)")
        V = self.num_clip_frames
        _, Cp, TV, Hh, Ww = panoptic_features.shape
        pf = panoptic_features.reshape(b, Cp, T, V, Hh, Ww).permute(0, 2, 1, 3, 4, 5).reshape(b * T, Cp, V * Hh, Ww).contiguous()   # CC:278
        x = clip_query.permute(0, 2, 1, 3).reshape(b, T * Q, C).contiguous().float()          # 'b q t c -> b (t q) c': rows (t, q)
        predictions_class, predictions_mask = [], []
        for i in range(self.num_layers):
            x = self.transformer_trajectory_self_attention_layers[i](x, seq_len=Q, num_frames=T)                 # LN(x + TA(x))
            x32, x16 = ops.cc_aspp_fwd(x.view(-1, C), self._packed_aspp(i, x.device), b, T, Q)                     # LN(ASPP(z) + z)
            x = x32.view(b, T * Q, C)
            # the mask branch (embedding projection -> mask head -> query x pixel contraction) runs fp32-grade (split precision): it
            # decides the per-pixel argmax labels, and its rows are few.  The class branch pools bf16 embeddings as before.
            ce = self._class_embedding_projection.run32(x32, torch.bfloat16)
            me = self._mask_embedding_projection.run32(x32)
            r = self._predictor(me, ce, pf, T, V)
            predictions_class.append(r["class_logits"])
            predictions_mask.append(r["mask_logits"])
        self.last_clip_query = x.view(b, T, Q, C).permute(0, 2, 1, 3)
        aux = []
        target_size = predictions_mask[-1].shape[-3:]
        align_corners = (target_size[-1] % 2 == 1)
        for a, m in zip(predictions_class[:-1], predictions_mask[:-1]):
            aux.append({"pred_logits": a, "pred_masks": torch.nn.functional.interpolate(m, size=target_size, mode="trilinear", align_corners=align_corners)})
        out = {"pred_logits": predictions_class[-1], "pred_masks": predictions_mask[-1], "aux_outputs": aux}
        if self.outputs_on_cpu:
            out = {"pred_logits": out["pred_logits"].cpu(), "pred_masks": out["pred_masks"].cpu(),
                   "aux_outputs": [{k: v.cpu() for k, v in a.items()} for a in aux]}
        return out

    def forward_sharded(self, clip_query_local, panoptic_local, n_clips, group=None, gather=None):
        """Clip-sharded video (one process per GPU): `clip_query_local` [1, Q, T_local, C] and
        `panoptic_local` [1, 128, T_local*V, H, W] hold this rank's contiguous clips; returns the FINAL layer's
        {'pred_logits', 'pred_masks'} for the whole video on every rank.  One all-gather of the clip queries before the
        layers (run redundantly), one of the mask logits after; see `sharding.cross_clip_sharded`."""
        from . import sharding
        _require_inference(self, clip_query_local, panoptic_local)
        b, Q, T_local, C = clip_query_local.shape
        if b != 1:
            raise NotImplementedError("axial_vs_b200: the cross-clip module runs one video at a time at inference (as the reference does)")
        V = self.num_clip_frames
        _, Cp, TV, Hh, Ww = panoptic_local.shape
        pf = panoptic_local.reshape(1, Cp, T_local, V, Hh, Ww).permute(0, 2, 1, 3, 4, 5).reshape(T_local, Cp, V * Hh, Ww).contiguous()

        def refine(cq_full):
            T = cq_full.shape[2]
            x = cq_full.permute(0, 2, 1, 3).reshape(1, T * Q, C).contiguous().float()
            x16 = None
            for i in range(self.num_layers):
                x = self.transformer_trajectory_self_attention_layers[i](x, seq_len=Q, num_frames=T)
                x32, x16 = ops.cc_aspp_fwd(x.view(-1, C), self._packed_aspp(i, x.device), 1, T, Q)
                x = x32.view(1, T * Q, C)
            ce = self._class_embedding_projection.run32(x32, torch.bfloat16)
            me = self._mask_embedding_projection.run32(x32)
            return self._predictor.class_logits(ce, T), self._predictor.mask_kernels(me)

        def masks(mk_local, pf_local, t_local):
            return self._predictor.mask_logits(mk_local, pf_local, t_local)

        cls, ml = sharding.cross_clip_sharded(refine, masks, clip_query_local.float(), pf, n_clips, group, gather)
        T = ml.shape[1]
        return {"pred_logits": cls, "pred_masks": ml.view(1, Q, T * V, Hh, Ww)}
