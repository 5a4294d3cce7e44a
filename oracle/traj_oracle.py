"""CPU oracle for the axial-trajectory attention hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-the-spec restatement of the reference's PyTorch modules, written with
plain tensor algebra (explicit per-frame loops, no einops) so that it is an independent check of
both the reference semantics and the CUDA kernels.  Runs in fp32 (default) or fp64 on the CPU.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this file, and only as the checker / the timed CPU arm.  The product package
(``axial_vs_b200``) never imports it and has no CPU fallback.

Parity status: PINNED.  The reference ships no golden vectors for this path, so the
pin is the reference itself: ``oracle/make_golden.py`` imports the unmodified reference modules
(``oracle/ref_loader.py``) in the authoring container and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this file against those fixtures (fp32, atol 2e-5), and
``tests/test_oracle_vs_reference.py`` re-runs the live comparison whenever ``/root/reference`` exists.

Reference citations use the prefixes of SURVEY.md:
  Vk/ = MaXTron_Video-kMaX/,  WC/ = Vk/maxtron_deeplab/modeling/within_clip_tracking_module/,
  CC  = Vk/maxtron_deeplab/modeling/cross_clip_tracking_module/maxtron_cross_clip_tracking_module.py

All ``params`` arguments are flat dicts keyed by the reference's own state-dict leaf names
(e.g. ``"height_attn.proj_kv.weight"``), so one state dict drives the reference, the oracle and
the CUDA modules.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------
def _sub(params: Params, prefix: str) -> Params:
    p = prefix + "."
    return {k[len(p):]: v for k, v in params.items() if k.startswith(p)}


# Error-budget emulation (tests/test_error_budget_cpu.py only): stages named here round their matmul operands to bf16 (fp32 /
# fp64 accumulation), which is what the CUDA path does; empty = the plain reference arithmetic.
#   "ta"     q/k/v/proj_q/proj_kv/proj/FFN linears and the attention operands (q, k, v, P, x, q2, o)
#   "aspp"   the temporal ASPP convolutions           "proj"   the ConvBN 1x1 projections / heads
#   "einsum" the query x pixel mask contraction
EMULATE_BF16: set = set()


def _rb(x: Tensor, stage: str) -> Tensor:
    return x.bfloat16().to(x.dtype) if stage in EMULATE_BF16 else x


def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    y = _rb(x, "ta") @ _rb(w.to(x.dtype), "ta").t()
    return y if b is None else y + b.to(x.dtype)


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm over the last dim, biased variance (torch default eps=1e-5)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w.to(x.dtype) + b.to(x.dtype)


def gelu(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def softmax_last(x: Tensor) -> Tensor:
    m = x.max(-1, keepdim=True).values
    e = torch.exp(x - m)
    return e / e.sum(-1, keepdim=True)


# --------------------------------------------------------------------------------------------
# A1 / A8: TrajectoryAttention            WC/temporal_attention.py:35-76,  CC:91-130
# --------------------------------------------------------------------------------------------
def trajectory_attention_core(q: Tensor, k: Tensor, v: Tensor, p: Params, num_frames: int,
                              num_heads: int = 8, return_maps: bool = False,
                              return_intermediates: bool = False):
    """Steps 2-8 of Appendix A given already-projected q, k, v of shape [B', N, C].

    q,k,v come from three Linear layers (WC/temporal_attention.py:42-44) or from one fused
    ``qkv`` Linear (CC:98).  ``p`` holds proj_q / proj_kv / proj weights and biases.
    """
    Bp, N, C = q.shape
    F = num_frames
    n = N // F
    assert n * F == N, "sequence length must be num_frames * tokens_per_frame"
    h = num_heads
    d = C // h
    scale = d ** -0.5                                        # :24-25

    def heads(t):                                            # 'b n (h d) -> b h n d'   :47-48
        return t.reshape(Bp, N, h, d).permute(0, 2, 1, 3)

    qh, kh, vh = heads(_rb(q, "ta")), heads(_rb(k, "ta")), heads(_rb(v, "ta"))   # [B', h, N, d]

    # spatial attention, softmax taken independently inside every key frame      :51-60
    x = q.new_zeros(Bp, N, F, C)
    maps = q.new_zeros(Bp, h, N, F, n) if return_maps else None
    for f in range(F):
        kf = kh[:, :, f * n:(f + 1) * n, :]                  # keys of frame f
        vf = vh[:, :, f * n:(f + 1) * n, :]
        s = scale * (qh @ kf.transpose(-1, -2))              # [B', h, N, n]
        a = softmax_last(s)                                  # softmax over the n keys of frame f :54
        if return_maps:
            maps[:, :, :, f, :] = a
        xf = _rb(a, "ta") @ vf                               # [B', h, N, d]           :56-57
        x[:, :, f, :] = xf.permute(0, 2, 1, 3).reshape(Bp, N, C)   # merge heads, head-major   :60

    # x_diag: for query token t = g*n + i take the aggregation over its own frame g      :61-63
    frame_of_token = torch.arange(N, device=q.device) // n
    x = _rb(x, "ta")
    x_diag = x[:, torch.arange(N, device=q.device), frame_of_token, :]        # [B', N, C]

    q2 = linear(x_diag, p["proj_q.weight"], p["proj_q.bias"]) * scale         # :64,67
    kv2 = linear(x, p["proj_kv.weight"], p["proj_kv.bias"])                   # [B', N, F, 2C]   :65
    k2, v2 = kv2[..., :C], kv2[..., C:]                                       # chunk(2): k2 first
    q2h = _rb(q2, "ta").reshape(Bp, N, h, d)
    k2h = k2.reshape(Bp, N, F, h, d)
    v2h = v2.reshape(Bp, N, F, h, d)
    logits = (q2h[:, :, None, :, :] * k2h).sum(-1)                            # [B', N, F, h]   :70
    a2 = softmax_last(logits.permute(0, 1, 3, 2))                             # softmax over F  :71
    o = (a2.permute(0, 1, 3, 2)[..., None] * v2h).sum(2)                      # [B', N, h, d]   :72
    o = _rb(o.reshape(Bp, N, C), "ta")                                        # :73
    y = linear(o, p["proj.weight"], p["proj.bias"])                           # :75

    if return_intermediates:
        return y, dict(x=x, x_diag=x_diag, q2=q2, kv2=kv2, o=o)
    if return_maps:
        # reference layout: [(B' h), N, F, n] with head the fast factor of the fused dim   :47,76
        return y, maps.reshape(Bp * h, N, F, n)
    return y, None


def trajectory_attention(query: Tensor, key: Tensor, value: Tensor, p: Params, num_frames: int = 2,
                         num_heads: int = 8, return_maps: bool = False):
    """Within-clip TrajectoryAttention.forward -- WC/temporal_attention.py:35-76."""
    q = linear(query, p["q.weight"], p["q.bias"])            # :42
    k = linear(key, p["k.weight"], p["k.bias"])              # :43
    v = linear(value, p["v.weight"], p["v.bias"])            # :44
    return trajectory_attention_core(q, k, v, p, num_frames, num_heads, return_maps)


def cc_trajectory_attention(x: Tensor, p: Params, seq_len: int = 128, num_frames: int = 6,
                            num_heads: int = 8) -> Tensor:
    """Cross-clip TrajectoryAttention.forward (fused qkv, no positional term) -- CC:91-130."""
    Bp, N, C = x.shape
    assert N == seq_len * num_frames
    qkv = linear(x, p["qkv.weight"], p["qkv.bias"])          # CC:98
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    return trajectory_attention_core(q, k, v, p, num_frames, num_heads)[0]


# --------------------------------------------------------------------------------------------
# A2 / A3 / A4: layers and encoder        WC/temporal_attention.py:79-220
# --------------------------------------------------------------------------------------------
def _ffn_tail(src: Tensor, p: Params, activation: str = "relu") -> Tensor:
    """norm1 -> linear1 -> act -> linear2 -> residual -> norm2   (:181-185, :217-218)."""
    src = layer_norm(src, p["norm1.weight"], p["norm1.bias"])
    hid = linear(src, p["linear1.weight"], p["linear1.bias"])
    hid = torch.relu(hid) if activation == "relu" else gelu(hid)
    src = src + linear(hid, p["linear2.weight"], p["linear2.bias"])
    return layer_norm(src, p["norm2.weight"], p["norm2.bias"])


def axial_layer(src: Tensor, pos: Tensor, p: Params, num_heads: int = 8, activation: str = "relu",
                return_maps: bool = False):
    """TemporalAxialTrajectoryAttentionLayer.forward -- WC/temporal_attention.py:187-220.

    src [B*T, H*W, C], pos [B, T, H, W, C] -> (src', h_map, w_map).
    """
    B, T, H, W, C = pos.shape
    s = src.reshape(B, T, H, W, C)
    # height pass: sequences (b, w), tokens (t, h)                                   :197-204
    sh = s.permute(0, 3, 1, 2, 4).reshape(B * W, T * H, C)
    ph = pos.permute(0, 3, 1, 2, 4).reshape(B * W, T * H, C)
    kq = sh + ph                                                                     # :200
    yh, hmap = trajectory_attention(kq, kq, sh, _sub(p, "height_attn"), T, num_heads, return_maps)
    sh = sh + yh                                                                     # :204
    s = sh.reshape(B, W, T, H, C).permute(0, 2, 3, 1, 4)                             # back to B T H W C
    # width pass: sequences (b, h), tokens (t, w)                                    :206-213
    sw = s.permute(0, 2, 1, 3, 4).reshape(B * H, T * W, C)
    pw = pos.permute(0, 2, 1, 3, 4).reshape(B * H, T * W, C)
    kq = sw + pw
    yw, wmap = trajectory_attention(kq, kq, sw, _sub(p, "width_attn"), T, num_heads, return_maps)
    sw = sw + yw
    s = sw.reshape(B, H, T, W, C).permute(0, 2, 1, 3, 4).reshape(B * T, H * W, C)    # :215
    return _ffn_tail(s, p, activation), hmap, wmap                                   # :217-218


def trajectory_layer(src: Tensor, pos: Tensor, p: Params, num_heads: int = 8, activation: str = "relu"):
    """TemporalTrajectoryAttentionLayer.forward (non-axial) -- WC/temporal_attention.py:131-155."""
    B, T = pos.shape[:2]
    C = src.shape[-1]
    s = src.reshape(B, -1, C)                                # '(B T) HW C -> B (T HW) C'   :141
    pp = pos.reshape(B, -1, C)                               # :142
    kq = s + pp
    y, _ = trajectory_attention(kq, kq, s, _sub(p, "temporal_attn"), T, num_heads)
    s = (s + y).reshape(src.shape)                           # :148-150
    return _ffn_tail(s, p, activation), None, None


def temporal_encoder(src: Tensor, pos: Tensor, layers: List[Params], attn_type: str = "axial-trajectory",
                     num_heads: int = 8, activation: str = "relu", return_maps: bool = False):
    """TemporalEncoder.forward -- WC/temporal_attention.py:90-100 (maps of the LAST layer only)."""
    hm = wm = None
    for p in layers:
        if attn_type == "axial-trajectory":
            src, hm, wm = axial_layer(src, pos, p, num_heads, activation, return_maps)
        elif attn_type == "trajectory":
            src, hm, wm = trajectory_layer(src, pos, p, num_heads, activation)
        else:
            raise ValueError(attn_type)
    return src, hm, wm


def split_encoder_params(state: Params, prefix: str = "temporal_layers") -> List[Params]:
    """Group a TemporalEncoder state dict ('temporal_layers.{i}.<leaf>') into per-layer dicts."""
    out: Dict[int, Params] = {}
    for k, v in state.items():
        if not k.startswith(prefix + "."):
            continue
        rest = k[len(prefix) + 1:]
        idx, leaf = rest.split(".", 1)
        out.setdefault(int(idx), {})[leaf] = v
    return [out[i] for i in sorted(out)]


# --------------------------------------------------------------------------------------------
# A5: 3-D sine positional table + level embed     WC/pos_embeddings.py:86-130, WC/msdeformattn.py:112-115
# --------------------------------------------------------------------------------------------
def pos3d_table(B: int, T: int, H: int, W: int, num_pos_feats: int = 128, temperature: float = 10000.0,
                normalize: bool = True, scale: float = 2 * math.pi, dtype=torch.float32) -> Tensor:
    """PositionEmbeddingSine3D.forward with mask=None, returned channels-last [B, T, H, W, 2*npf]."""
    npf = num_pos_feats
    z = torch.arange(1, T + 1, dtype=torch.float32)          # cumsum of ones     :99-102
    y = torch.arange(1, H + 1, dtype=torch.float32)
    x = torch.arange(1, W + 1, dtype=torch.float32)
    if normalize:                                            # :103-107
        eps = 1e-6
        z = z / (z[-1] + eps) * scale
        y = y / (y[-1] + eps) * scale
        x = x / (x[-1] + eps) * scale
    i = torch.arange(npf, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="trunc") / npf)            # :109-111
    iz = torch.arange(2 * npf, dtype=torch.float32)
    dim_tz = temperature ** (2 * torch.div(iz, 2, rounding_mode="trunc") / (2 * npf))    # :113-115

    def interleave(arg):                                     # even ch -> sin, odd ch -> cos   :120-122
        out = torch.empty_like(arg)
        out[..., 0::2] = arg[..., 0::2].sin()
        out[..., 1::2] = arg[..., 1::2].cos()
        return out

    px = interleave(x[:, None] / dim_t)                      # [W, npf]
    py = interleave(y[:, None] / dim_t)                      # [H, npf]
    pz = interleave(z[:, None] / dim_tz)                     # [T, 2npf]
    pos = torch.empty(T, H, W, 2 * npf, dtype=torch.float32)
    pos[..., :npf] = py[None, :, None, :]                    # cat((pos_y, pos_x))      :123
    pos[..., npf:] = px[None, None, :, :]
    pos = pos + pz[:, None, None, :]
    return pos[None].expand(B, T, H, W, 2 * npf).contiguous().to(dtype)


def level_pos3d(B: int, T: int, H: int, W: int, level_embed: Tensor, **kw) -> Tensor:
    """pos_3d of one level as the layer receives it: table + level_embed_3d[lvl]  (WC/msdeformattn.py:112-115)."""
    return pos3d_table(B, T, H, W, **kw) + level_embed.reshape(1, 1, 1, 1, -1).to(kw.get("dtype", torch.float32))


# --------------------------------------------------------------------------------------------
# A9 / A10: cross-clip layer, ASPP, predictor       CC:30-75, 133-201, 275-322
# --------------------------------------------------------------------------------------------
def cc_attention_layer(x: Tensor, p: Params, seq_len: int, num_frames: int, num_heads: int = 8) -> Tensor:
    """TrajectoryAttentionLayer.forward_post -- CC:156-161 (normalize_before=False everywhere)."""
    y = cc_trajectory_attention(x, _sub(p, "self_attn"), seq_len, num_frames, num_heads)
    return layer_norm(x + y, p["norm.weight"], p["norm.bias"])


def conv1d_same_replicate(x: Tensor, w: Tensor, b: Optional[Tensor], dilation: int) -> Tensor:
    """nn.Conv1d(k, stride 1, dilation, padding='same', padding_mode='replicate') -- CC:180-182.

    x [M, Cin, T], w [Cout, Cin, k].  'same' padding for odd k is dilation*(k-1)/2 each side; the
    replicate mode clamps the time index.
    """
    M, Cin, T = x.shape
    Cout, _, k = w.shape
    total = dilation * (k - 1)
    left = total // 2
    out = x.new_zeros(M, Cout, T)
    t_idx = torch.arange(T)
    for j in range(k):
        src_t = (t_idx - left + j * dilation).clamp(0, T - 1)
        out += torch.einsum("oc,mct->mot", _rb(w[:, :, j].to(x.dtype), "aspp"), _rb(x[:, :, src_t], "aspp"))
    if b is not None:
        out = out + b.to(x.dtype)[None, :, None]
    return out


def layer_norm_channels_first(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """convnext.LayerNorm(data_format='channels_first') on [M, C, T] -- Vk/kmax_deeplab/modeling/backbone/convnext.py:52-80."""
    u = x.mean(1, keepdim=True)
    s = ((x - u) ** 2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w.to(x.dtype)[:, None] * x + b.to(x.dtype)[:, None]


def batch_norm_eval(x: Tensor, p: Params, prefix: str, eps: float = 1e-3) -> Tensor:
    """nn.SyncBatchNorm(eps=1e-3) in eval mode == affine with running stats (channel dim 1)."""
    shape = [1, -1] + [1] * (x.dim() - 2)
    rm, rv = p[prefix + ".running_mean"].to(x.dtype), p[prefix + ".running_var"].to(x.dtype)
    g, bt = p[prefix + ".weight"].to(x.dtype), p[prefix + ".bias"].to(x.dtype)
    return (x - rm.reshape(shape)) / torch.sqrt(rv.reshape(shape) + eps) * g.reshape(shape) + bt.reshape(shape)


def aspp(x: Tensor, p: Params, atrous_rates=(1, 2, 3), norm_fn: str = "ln") -> Tensor:
    """ASPP.forward -- CC:190-201.  x [M, C, T]."""
    r = [conv1d_same_replicate(x, p[f"_aspp_conv{i}.weight"], p[f"_aspp_conv{i}.bias"], atrous_rates[i])
         for i in range(3)]
    z = torch.cat(r, dim=1)                                                       # [M, 3C, T]
    z = torch.einsum("oc,mct->mot", _rb(p["_proj_conv_bn_act.conv.weight"][:, :, 0].to(x.dtype), "aspp"), _rb(z, "aspp"))   # 1x1, no bias
    if norm_fn == "ln":
        z = layer_norm_channels_first(z, p["_proj_conv_bn_act.norm.weight"], p["_proj_conv_bn_act.norm.bias"])
    elif norm_fn == "syncbn":
        z = batch_norm_eval(z, p, "_proj_conv_bn_act.norm")
    return gelu(z)                                                                # dropout identity in eval


def conv_bn_1d(x: Tensor, p: Params, prefix: str, norm: Optional[str], act: Optional[str]) -> Tensor:
    """ConvBN(conv_type='1d', kernel_size=1) in eval -- Vk/kmax_deeplab/modeling/pixel_decoder/kmax_pixel_decoder.py:42-72."""
    w = p[prefix + ".conv.weight"][:, :, 0].to(x.dtype)
    y = torch.einsum("oc,bcn->bon", _rb(w, "proj"), _rb(x, "proj"))
    if prefix + ".conv.bias" in p:
        y = y + p[prefix + ".conv.bias"].to(x.dtype)[None, :, None]
    if norm == "syncbn":
        y = batch_norm_eval(y, p, prefix + ".norm")
    if act == "gelu":
        y = gelu(y)
    return y


def cc_predictor(mask_emb: Tensor, class_emb: Tensor, pixel_feature: Tensor, p: Params,
                 num_clips: int, num_clip_frames: int) -> Dict[str, Tensor]:
    """MaXTronCCPredictor.forward, eval branch -- CC:45-75.

    mask_emb/class_emb [T', C, Q]; pixel_feature [T', 128, V*H, W] (batch of 1 video).
    """
    act = conv_bn_1d(class_emb, p, "_transformer_class_activation_head", None, None)      # [T',1,Q]
    act = torch.softmax(act, dim=0)                                                        # over clips :48-49
    ce = (class_emb * act).sum(0, keepdim=True)                                            # :50
    cls = conv_bn_1d(ce, p, "_transformer_class_head", None, None).permute(0, 2, 1)        # [1,Q,K]  :51
    bias = torch.zeros(cls.shape[-1], dtype=cls.dtype)
    bias[-1] = math.log((cls.shape[-1] - 1) * 0.9 / (1 - 0.9))                             # add_bias_towards_void
    cls = cls + bias
    mk = conv_bn_1d(mask_emb, p, "_transformer_mask_head", "syncbn", None)                 # [T',128,Q] :53
    logits = torch.einsum("bchw,bcn->bnhw", _rb(pixel_feature, "einsum"), _rb(mk, "einsum"))   # :62-67
    logits = batch_norm_eval(logits.unsqueeze(1), p, "_pixel_space_mask_batch_norm").squeeze(1)   # :68
    Tp, Q, VH, Wd = logits.shape
    V = num_clip_frames
    # '(B T) C (V H) W -> B C (T V) H W'                                                    :69
    logits = logits.reshape(1, Tp, Q, V, VH // V, Wd).permute(0, 2, 1, 3, 4, 5).reshape(1, Q, Tp * V, VH // V, Wd)
    return {"class_logits": cls, "mask_logits": logits}


def cross_clip_module(clip_query: Tensor, panoptic_features: Tensor, p: Params, num_layers: int,
                      num_clip_frames: int, atrous_rates=(1, 2, 3), norm_fn: str = "ln",
                      num_heads: int = 8) -> Dict[str, Tensor]:
    """CrossClipTrackingModule.forward (final-layer outputs) -- CC:275-322.

    clip_query [b, Q, T, C]; panoptic_features [b, 128, T*V, H, W].
    """
    b, Q, T, C = clip_query.shape
    V = num_clip_frames
    _, Cp, TV, Hh, Ww = panoptic_features.shape
    # 'B C (T V) H W -> (B T) C (V H) W'                                                     :278
    pf = panoptic_features.reshape(b, Cp, T, V, Hh, Ww).permute(0, 2, 1, 3, 4, 5).reshape(b * T, Cp, V * Hh, Ww)
    outs = []
    cq = clip_query
    for i in range(num_layers):
        x = cq.permute(0, 2, 1, 3).reshape(b, T * Q, C)                                   # 'b q t c -> b (t q) c' :284
        x = cc_attention_layer(x, _sub(p, f"transformer_trajectory_self_attention_layers.{i}"), Q, T, num_heads)
        z = x.reshape(b, T, Q, C).permute(0, 2, 3, 1).reshape(b * Q, C, T)                # 'b (t q) c -> (b q) c t' :290
        z = aspp(z, _sub(p, f"conv_short_aggregate_layers.{i}"), atrous_rates, norm_fn) + z
        z = layer_norm(z.transpose(1, 2), p[f"conv_norms.{i}.weight"], p[f"conv_norms.{i}.bias"])   # :293-295
        cq = z.reshape(b, Q, T, C)
        vq = cq.permute(0, 2, 3, 1).reshape(b * T, C, Q)                                  # 'b q t c -> (b t) c q'
        ce = conv_bn_1d(vq, p, "_class_embedding_projection", "syncbn", "gelu")
        me = conv_bn_1d(vq, p, "_mask_embedding_projection", "syncbn", "gelu")
        outs.append(cc_predictor(me, ce, pf, _sub(p, "_predictor"), T, V))
    return {"pred_logits": outs[-1]["class_logits"], "pred_masks": outs[-1]["mask_logits"],
            "clip_query": cq, "all": outs}


# --------------------------------------------------------------------------------------------
# algorithmic work -- shared by bench.py and the tests
# --------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------ clip-level decoder attention (row A11)
# DEC = Vk/maxtron_deeplab/modeling/transformer_decoder/maxtron_transformer_decoder.py
def query_self_attention(q: Tensor, k: Tensor, v: Tensor, p: Params) -> Tensor:
    """AttentionOperation.forward, eval mode -- DEC:56-71.  q, k [N,h,dk,L]; v [N,h,dv,L] -> [N, h*dv, L]."""
    N, h, dv, L = v.shape
    sim = torch.einsum("bhdl,bhdm->bhlm", q, k)                                   # :58
    sim = batch_norm_eval(sim, p, "_batch_norm_similarity")                       # :59  (per-head BN on the logits)
    w = torch.softmax(sim, dim=-1)                                                # :61-62
    r = torch.einsum("bhlm,bhdm->bhdl", w, v).reshape(N, h * dv, L)               # :63-65
    r = batch_norm_eval(r, p, "_batch_norm_retrieved_value")                      # :66-67
    return gelu(r)                                                                # :68


def kmeans_update(mask_logits: Tensor, pixel_value: Tensor, advanced: bool = False):
    """k-means assignment + update of kMaXTransformerLayer.forward -- DEC:196-208.

    mask_logits [N, L, M], pixel_value [N, D, M] -> (update [N, D, L], assignment [N, M]).  The assignment is the first
    maximum over the L cluster centres (torch.max semantics on CPU).
    """
    N, L, M = mask_logits.shape
    index = mask_logits.max(1, keepdim=True)[1]                                   # :199
    onehot = torch.zeros_like(mask_logits).scatter_(1, index, 1.0)                # :200
    upd = torch.einsum("blm,bdm->bdl", onehot, pixel_value)                       # :204
    if advanced:
        upd = upd / torch.clamp(onehot.sum(-1).unsqueeze(1), min=1.0)             # :206-208
    return upd, index[:, 0]


# ------------------------------------------------------------------------------------------------ within-clip projections (row f1)
def group_norm(x: Tensor, num_groups: int, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.GroupNorm on [images, C, ...]: statistics per image over (C/groups channels x all positions), biased variance."""
    n, C = x.shape[0], x.shape[1]
    xg = x.reshape(n, num_groups, -1)
    mu = xg.mean(-1, keepdim=True)
    var = ((xg - mu) ** 2).mean(-1, keepdim=True)
    y = ((xg - mu) / torch.sqrt(var + eps)).reshape(x.shape)
    shape = [1, C] + [1] * (x.dim() - 2)
    return y * w.to(x.dtype).reshape(shape) + b.to(x.dtype).reshape(shape)


def input_proj(x: Tensor, p: Params) -> Tensor:
    """input_proj[idx](x) then flatten(2).transpose(1, 2) -- WC/msdeformattn.py:355-358, 413-416 and :100-106.
    x [images, c_in, H, W] -> tokens [images, H*W, 256]."""
    y = torch.einsum("oc,nchw->nohw", p["0.weight"][:, :, 0, 0].to(x.dtype), x) + p["0.bias"].to(x.dtype)[None, :, None, None]
    y = group_norm(y, 32, p["1.weight"], p["1.bias"])
    return y.flatten(2).transpose(1, 2)


def output_proj(tokens: Tensor, p: Params, H: int, W: int) -> Tensor:
    """output_proj[i](z.transpose(1, 2).view(bs, -1, H, W)) -- WC/msdeformattn.py:359-362, 432-434.
    tokens [images, H*W, 256] -> [images, c_out, H, W]."""
    z = tokens.transpose(1, 2).reshape(tokens.shape[0], tokens.shape[2], H, W)
    y = torch.einsum("oc,nchw->nohw", p["0.weight"][:, :, 0, 0].to(z.dtype), z) + p["0.bias"].to(z.dtype)[None, :, None, None]
    return group_norm(y, 32, p["1.weight"], p["1.bias"])


# ------------------------------------------------------------------------------------------------ MSDeformAttn spatial layer (row f2)
# MSDA = WC/ops/modules/ms_deform_attn.py, CORE = WC/ops/functions/ms_deform_attn_func.py, ENC = WC/msdeformattn.py
def msda_reference_points(spatial_shapes, n_images: int) -> Tensor:
    """ENC:231-245 with valid_ratios == 1 (the module builds all-False masks, ENC:92): the centre of every token's own cell,
    normalised to [0,1], repeated for every level.  -> [n_images, sum(H*W), n_levels, 2] as (x, y)."""
    pts = []
    for (H, W) in spatial_shapes:
        ys, xs = torch.meshgrid(torch.linspace(0.5, H - 0.5, H), torch.linspace(0.5, W - 0.5, W), indexing="ij")
        pts.append(torch.stack((xs.reshape(-1) / W, ys.reshape(-1) / H), -1))
    ref = torch.cat(pts, 0)                                                     # [Len, 2]
    return ref[None, :, None, :].expand(n_images, -1, len(spatial_shapes), -1).contiguous()


def ms_deform_attn(query: Tensor, reference_points: Tensor, src: Tensor, spatial_shapes, p: Params,
                   n_heads: int = 8, n_points: int = 4, sample_only: bool = False) -> Tensor:
    """MSDeformAttn.forward without padding mask -- MSDA:92-125 and the bilinear sampling of CORE:51-72
    (grid_sample, align_corners=False, zero padding), written as explicit gathers."""
    N, Lq, C = query.shape
    L = len(spatial_shapes)
    d = C // n_heads
    value = linear(src, p["value_proj.weight"], p["value_proj.bias"]).reshape(N, -1, n_heads, d)          # :98-101
    off = linear(query, p["sampling_offsets.weight"], p["sampling_offsets.bias"]).reshape(N, Lq, n_heads, L, n_points, 2)
    aw = linear(query, p["attention_weights.weight"], p["attention_weights.bias"]).reshape(N, Lq, n_heads, L * n_points)
    aw = torch.softmax(aw, -1).reshape(N, Lq, n_heads, L, n_points)                                       # :104
    out = query.new_zeros(N, Lq, n_heads, d)
    start = 0
    for l, (H, W) in enumerate(spatial_shapes):
        v = value[:, start:start + H * W]                                                                 # [N, H*W, h, d]
        start += H * W
        norm = torch.tensor([W, H], dtype=query.dtype)
        loc = reference_points[:, :, None, l, None, :] + off[:, :, :, l] / norm                           # :107-109  [N,Lq,h,P,2]
        x = loc[..., 0] * W - 0.5                                                                         # align_corners=False
        y = loc[..., 1] * H - 0.5
        x0, y0 = torch.floor(x), torch.floor(y)
        fx, fy = x - x0, y - y0
        for dy, wy in ((0, 1 - fy), (1, fy)):
            for dx, wx in ((0, 1 - fx), (1, fx)):
                xi, yi = (x0 + dx).long(), (y0 + dy).long()
                ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)                                          # zero padding
                idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1))                                       # [N,Lq,h,P]
                g = torch.gather(v.permute(0, 2, 1, 3), 2,                                                # [N,h,HW,d] gathered at [N,h,Lq*P]
                                 idx.permute(0, 2, 1, 3).reshape(N, n_heads, -1, 1).expand(-1, -1, -1, d))
                g = g.reshape(N, n_heads, Lq, n_points, d).permute(0, 2, 1, 3, 4)                         # [N,Lq,h,P,d]
                wgt = (wx * wy * ok.to(query.dtype) * aw[:, :, :, l])[..., None]
                out = out + (g * wgt).sum(3)
    if sample_only:                                                                                       # the Tube-Link module works on this
        return out.reshape(N, Lq, C)
    return linear(out.reshape(N, Lq, C), p["output_proj.weight"], p["output_proj.bias"])                  # :124


def tl_axial_trajectory_msda(query: Tensor, query_pos: Tensor, query_pos3d: List[Tensor], reference_points: Tensor, spatial_shapes,
                             p: Params, num_temporal_levels: int = 2, n_points: int = 4, identity: Optional[Tensor] = None) -> Tensor:
    """MultiScaleDeformableAxialTrajectoryAttention.forward (batch_first, skip_connect, eval) --
    TL/mmdet/models/plugins/msdeformattn_pixel_decoder.py:556-638.  query [bs, num_query, C] over all levels; query_pos3d[i] [B, T, H_i, W_i, C]
    with bs = B*T.  `p`: sampling_offsets / attention_weights / value_proj / output_proj, `temporal_layer.temporal_layers.{k}.*`, `gamma`."""
    identity = query if identity is None else identity                                                    # :558-561
    sampled = ms_deform_attn(query + query_pos, reference_points, query, spatial_shapes, p, n_points=n_points, sample_only=True)   # :563-614
    sizes = [h * w for h, w in spatial_shapes]
    outs = list(torch.split(sampled, sizes, dim=1))                                                       # :618-620
    layers = split_encoder_params(_sub(p, "temporal_layer"))
    for i in range(num_temporal_levels):                                                                  # :622-627
        f = outs[i]
        t = f
        for lp in layers:                                                                                 # TL TemporalEncoder :724-727 (features only)
            t = axial_layer(t, query_pos3d[i], lp)[0]
        outs[i] = f + p["gamma"].to(f.dtype) * t
    out = linear(torch.cat(outs, dim=1), p["output_proj.weight"], p["output_proj.bias"])                  # :630-632
    return out + identity                                                                                 # :638 (dropout = identity in eval)


def tl_forward_head_clips(decoder_out: Tensor, mask_feature: Tensor, p: Params):
    """Mask2FormerVideoCCHeadTube.forward_head_clips + pred_class -- TL/models/video/tube_link_vis/mask2former_video_cc_head.py:761-797.
    decoder_out [t, l, q, b, c] (clips, layers, queries, batch, channels); mask_feature [b, T_frames, c_m, h, w].
    `p`: post_norm.{weight,bias}, activation_proj, cls_embed, mask_embed.{0,2,4} (Linear-ReLU-Linear-ReLU-Linear).
    Returns (class logits [l, b, q, K+1], mask logits [l, b, T_frames, q, h, w])."""
    num_clips = decoder_out.shape[0]
    fpc = mask_feature.shape[1] // num_clips
    x = layer_norm(decoder_out, p["post_norm.weight"], p["post_norm.bias"])                               # :768
    x = x.permute(1, 3, 0, 2, 4)                                                                          # (l, b, t, q, c)  :769
    act = torch.softmax(linear(x, p["activation_proj.weight"], p["activation_proj.bias"]), dim=2)        # :790
    cls = linear((x * act).sum(dim=2), p["cls_embed.weight"], p["cls_embed.bias"])                        # :791-796
    me = linear(torch.relu(linear(torch.relu(linear(x, p["mask_embed.0.weight"], p["mask_embed.0.bias"])),
                                  p["mask_embed.2.weight"], p["mask_embed.2.bias"])), p["mask_embed.4.weight"], p["mask_embed.4.bias"])   # :772
    masks = [torch.einsum("lbqc,btchw->lbtqhw", me[:, :, k], mask_feature[:, fpc * k:fpc * (k + 1)]) for k in range(num_clips)]   # :774-777
    return cls, torch.cat(masks, dim=2)


def msda_encoder_layer(src: Tensor, pos: Tensor, reference_points: Tensor, spatial_shapes, p: Params) -> Tensor:
    """MSDeformAttnTransformerEncoderLayer.forward (eval) -- ENC:205-215: src = LN1(src + MSDA(src+pos, ref, src)); LN2(src + FFN(src))."""
    src2 = ms_deform_attn(src + pos, reference_points, src, spatial_shapes, _sub(p, "self_attn"))
    src = layer_norm(src + src2, p["norm1.weight"], p["norm1.bias"])
    return _ffn_tail_only(src, p)


def _ffn_tail_only(src: Tensor, p: Params) -> Tensor:
    h = torch.relu(linear(src, p["linear1.weight"], p["linear1.bias"]))
    return layer_norm(src + linear(h, p["linear2.weight"], p["linear2.bias"]), p["norm2.weight"], p["norm2.bias"])


def within_clip_encoder(src: Tensor, spatial_shapes, pos: Tensor, pos_3d: List[Tensor], spatial: List[Params],
                        temporal: List[List[Params]], num_temporal_levels: int) -> Tensor:
    """MSDeformAttnTransformerEncoder.forward (unpadded) -- ENC:247-273: per stage the spatial layer on every level, then the
    stage's TemporalEncoder on the first `num_temporal_levels` levels (split / concat along the token axis, ENC:251-266)."""
    ref = msda_reference_points(spatial_shapes, src.shape[0])
    sizes = [h * w for h, w in spatial_shapes]
    out = src
    for sp, tl in zip(spatial, temporal):
        out = msda_encoder_layer(out, pos, ref, spatial_shapes, sp)
        parts = list(torch.split(out, sizes, dim=1))
        for i in range(num_temporal_levels):
            parts[i], _, _ = temporal_encoder(parts[i], pos_3d[i], tl, "axial-trajectory")
        out = torch.cat(parts, dim=1)
    return out


def pos2d_table(H: int, W: int, num_pos_feats: int = 128, temperature: float = 10000.0) -> Tensor:
    """PositionEmbeddingSine(normalize=True) without mask -- WC/pos_embeddings.py:30-53 -> [H*W, 2*num_pos_feats] (pos_y | pos_x)."""
    scale = 2 * math.pi
    y = torch.arange(1, H + 1, dtype=torch.float32) / (H + 1e-6) * scale
    x = torch.arange(1, W + 1, dtype=torch.float32) / (W + 1e-6) * scale
    i = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="trunc") / num_pos_feats)

    def enc(v):
        a = v[:, None] / dim_t
        return torch.stack((a[:, 0::2].sin(), a[:, 1::2].cos()), dim=2).flatten(1)

    py, px = enc(y), enc(x)
    return torch.cat((py[:, None, :].expand(H, W, -1), px[None, :, :].expand(H, W, -1)), -1).reshape(H * W, 2 * num_pos_feats)


def within_clip_module(features: List[Tensor], p: Params, B: int, T: int, num_stages: int = 2, num_temporal_levels: int = 2) -> List[Tensor]:
    """MSDeformAttnPixelDecoder.forward_features -- ENC:404-435 with MSDeformAttnTransformerEncoderOnly.forward ENC:91-125.
    `features` top-down (res5, res4, res3), each [B*T, C_l, H_l, W_l]; returns the output-projected maps in the same order."""
    shapes = [(int(f.shape[2]), int(f.shape[3])) for f in features]
    toks = [input_proj(f, _sub(p, f"input_proj.{i}")) for i, f in enumerate(features)]
    pos = torch.cat([pos2d_table(h, w) + p["transformer.level_embed_2d"][i] for i, (h, w) in enumerate(shapes)], 0)
    pos = pos[None].expand(B * T, -1, -1)
    pos3d = [level_pos3d(B, T, h, w, p["transformer.level_embed_3d"][i]) for i, (h, w) in enumerate(shapes[:num_temporal_levels])]
    enc = _sub(p, "transformer.encoder")
    spatial = [_sub(enc, f"spatial_layers.{i}") for i in range(num_stages)]
    temporal = [split_encoder_params(_sub(enc, f"temporal_layers.{i}")) for i in range(num_stages)]
    y = within_clip_encoder(torch.cat(toks, 1), shapes, pos, pos3d, spatial, temporal, num_temporal_levels)
    outs, start = [], 0
    for i, (h, w) in enumerate(shapes):
        outs.append(output_proj(y[:, start:start + h * w], _sub(p, f"output_proj.{i}"), h, w))
        start += h * w
    return outs


def flops_trajectory_attention(Bp: int, N: int, F: int, C: int = 256) -> int:
    return Bp * N * C * (10 * C + 4 * F * C + 4 * N + 4 * F)


def flops_axial_layer(B: int, T: int, H: int, W: int, C: int = 256, d_ffn: int = 1024) -> int:
    tokens = B * T * H * W
    return tokens * C * (20 * C + 8 * T * C + 4 * T * (H + W) + 8 * T + 4 * d_ffn)
