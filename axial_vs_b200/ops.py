"""Host-side operators over the C ABI (include/axvs.h): tensor checks, output/workspace allocation with the
caching allocator, current-stream plumbing, and `torch.library` registration (`torch.ops.axialvs.*`).

PyTorch is used for device memory and streams only; all arithmetic runs in libaxvs.so.  CPU tensors are
rejected with a RuntimeError -- there is no CPU fallback (BASELINE.json north_star; same behaviour as the
reference's native op, WC/ops/src/ms_deform_attn.h:43 `AT_ERROR("Not implemented on the CPU")`).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import AXIS_H, AXIS_NONE, AXIS_W, LayerWeights, TaWeights  # noqa: F401

C = 256          # d_model the kernels are specialised for
HEADS = 8
HEAD_DIM = 32


# ------------------------------------------------------------------------------------------------ helpers
def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _check(t: torch.Tensor, name: str, dtype: torch.dtype, shape: Optional[Sequence[int]] = None) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: axial_vs_b200 has no CPU implementation")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")


_workspaces: "Dict[Tuple[int, int], torch.Tensor]" = {}     # insertion-ordered: least recently used first
MAX_WORKSPACES = 8                                           # per process; callers cycling through many streams evict the oldest


def workspace(nbytes: int, dev: torch.device) -> torch.Tensor:
    """Per-(device, stream) scratch buffer, grown on demand; reuse is ordered by the stream itself.  At most MAX_WORKSPACES buffers
    are kept (least recently used evicted; an evicted buffer's memory returns to the caching allocator once its stream's work is
    done, which `record_stream` guarantees)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), _stream(dev))
    buf = _workspaces.pop(key, None)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        buf.record_stream(torch.cuda.current_stream(dev))
    _workspaces[key] = buf                                   # most recently used last
    while len(_workspaces) > MAX_WORKSPACES:
        _workspaces.pop(next(iter(_workspaces)))
    return buf


def release_workspaces() -> None:
    """Drop every cached scratch buffer (e.g. between a benchmark's configurations)."""
    _workspaces.clear()


# ------------------------------------------------------------------------------------------------ weights
def pack_weight(w: torch.Tensor) -> torch.Tensor:
    """fp32 nn.Linear weight [n_out, k] -> bf16 UMMA shared-memory image (uint8 tensor)."""
    _check(w, "weight", torch.float32)
    n_out, k = w.shape
    lib = _lib.load()
    out = torch.empty(lib.axvs_packed_weight_bytes(n_out, k), dtype=torch.uint8, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(lib.axvs_pack_weight(w.data_ptr(), n_out, k, out.data_ptr(), _stream(w.device)), "axvs_pack_weight")
    return out


def pack_weight_units(w: torch.Tensor, k_major: int = 0) -> torch.Tensor:
    """fp32 weight [n_out, k] -> 32 KiB-unit image of the fused kernels (see include/axvs.h)."""
    _check(w, "weight", torch.float32)
    n_out, k = w.shape
    lib = _lib.load()
    out = torch.empty(lib.axvs_packed_weight_bytes(n_out, k), dtype=torch.uint8, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(lib.axvs_pack_weight_units(w.data_ptr(), n_out, k, int(k_major), out.data_ptr(), _stream(w.device)),
                   "axvs_pack_weight_units")
    return out


@dataclass
class PackedTA:
    """Device-resident packed parameters of one TrajectoryAttention (struct axvs_ta_weights)."""
    w_qkv: torch.Tensor
    b_qkv: torch.Tensor
    w_pq: torch.Tensor
    b_pq: torch.Tensor
    w_pkv: torch.Tensor
    b_pkv: torch.Tensor
    w_qkv_u: torch.Tensor
    w_pq_u: torch.Tensor
    w_pkv_u: torch.Tensor
    w_proj_u: torch.Tensor
    w_proj: torch.Tensor
    b_proj: torch.Tensor

    def tensors(self) -> List[torch.Tensor]:
        return [self.w_qkv, self.b_qkv, self.w_pq, self.b_pq, self.w_pkv, self.b_pkv, self.w_qkv_u, self.w_pq_u, self.w_pkv_u,
                self.w_proj_u, self.w_proj, self.b_proj]

    def struct(self) -> TaWeights:
        return TaWeights(*[t.data_ptr() for t in self.tensors()])

    @staticmethod
    def from_tensors(ts: Sequence[torch.Tensor]) -> "PackedTA":
        return PackedTA(*ts)


def pack_ta(p: Dict[str, torch.Tensor], prefix: str = "") -> PackedTA:
    """Pack a TrajectoryAttention state dict (leaf names q/k/v or qkv, proj_q, proj_kv, proj)."""
    def g(name):
        return p[prefix + name].detach().float().contiguous()

    if prefix + "qkv.weight" in p:
        wqkv, bqkv = g("qkv.weight"), g("qkv.bias")
    else:
        wqkv = torch.cat([g("q.weight"), g("k.weight"), g("v.weight")], 0).contiguous()
        bqkv = torch.cat([g("q.bias"), g("k.bias"), g("v.bias")], 0).contiguous()
    wkv = g("proj_kv.weight")
    # rows re-ordered per head pair c: [k2 rows 64c..64c+63 ; v2 rows 256+64c..256+64c+63]  (fused trajectory kernel)
    order = torch.cat([torch.cat([torch.arange(64 * c, 64 * c + 64), torch.arange(256 + 64 * c, 256 + 64 * c + 64)]) for c in range(4)])
    wkv_c = wkv[order.to(wkv.device)].contiguous()
    return PackedTA(pack_weight(wqkv), bqkv, pack_weight(g("proj_q.weight")), g("proj_q.bias"),
                    pack_weight(wkv), g("proj_kv.bias"), pack_weight_units(wqkv), pack_weight_units(g("proj_q.weight")), pack_weight_units(wkv_c),
                    pack_weight_units(g("proj.weight")), pack_weight(g("proj.weight")), g("proj.bias"))


@dataclass
class PackedLayer:
    """struct axvs_layer_weights."""
    attn_h: PackedTA
    attn_w: Optional[PackedTA]
    ln1_g: torch.Tensor
    ln1_b: torch.Tensor
    w_ffn1: torch.Tensor
    b_ffn1: torch.Tensor
    w_ffn2: torch.Tensor
    b_ffn2: torch.Tensor
    w_ffn1_u: torch.Tensor
    w_ffn2_u: torch.Tensor
    ln2_g: torch.Tensor
    ln2_b: torch.Tensor
    d_ffn: int
    w_ffn1_n: Optional[torch.Tensor] = None      # linear1 as N = 256 units (pack_weight_units mode 2)

    def tensors(self) -> List[torch.Tensor]:
        aw = self.attn_w if self.attn_w is not None else self.attn_h
        return self.attn_h.tensors() + aw.tensors() + [self.ln1_g, self.ln1_b, self.w_ffn1, self.b_ffn1, self.w_ffn2,
                                                       self.b_ffn2, self.w_ffn1_u, self.w_ffn2_u, self.ln2_g, self.ln2_b, self.w_ffn1_n]

    def struct(self) -> LayerWeights:
        aw = self.attn_w if self.attn_w is not None else self.attn_h
        return LayerWeights(self.attn_h.struct(), aw.struct(), self.ln1_g.data_ptr(), self.ln1_b.data_ptr(),
                            self.w_ffn1.data_ptr(), self.b_ffn1.data_ptr(), self.w_ffn2.data_ptr(), self.b_ffn2.data_ptr(),
                            self.w_ffn1_u.data_ptr(), self.w_ffn2_u.data_ptr(), self.ln2_g.data_ptr(), self.ln2_b.data_ptr(), self.d_ffn,
                            self.w_ffn1_n.data_ptr() if self.w_ffn1_n is not None else None)

    @staticmethod
    def from_tensors(ts: Sequence[torch.Tensor], d_ffn: int) -> "PackedLayer":
        return PackedLayer(PackedTA.from_tensors(ts[0:12]), PackedTA.from_tensors(ts[12:24]), *ts[24:34], d_ffn=d_ffn,
                           w_ffn1_n=ts[34] if len(ts) > 34 else None)


def pack_layer(p: Dict[str, torch.Tensor], axial: bool = True) -> PackedLayer:
    """Pack the state dict of a Temporal(Axial)TrajectoryAttentionLayer (WC/temporal_attention.py:159-175)."""
    def g(name):
        return p[name].detach().float().contiguous()

    ah = pack_ta(p, "height_attn." if axial else "temporal_attn.")
    aw = pack_ta(p, "width_attn.") if axial else None
    return PackedLayer(ah, aw, g("norm1.weight"), g("norm1.bias"), pack_weight(g("linear1.weight")), g("linear1.bias"),
                       pack_weight(g("linear2.weight")), g("linear2.bias"), pack_weight_units(g("linear1.weight")),
                       pack_weight_units(g("linear2.weight"), k_major=1), g("norm2.weight"), g("norm2.bias"),
                       d_ffn=p["linear1.weight"].shape[0], w_ffn1_n=pack_weight_units(g("linear1.weight"), k_major=2))


# ------------------------------------------------------------------------------------------------ operators
def linear(a: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], n_out: int, *, scale: float = 1.0,
           relu: bool = False, out_dtype: torch.dtype = torch.bfloat16, resid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = act((a @ W^T + bias) * scale) (+ resid): tcgen05 GEMM.  a bf16 [M, K]."""
    _check(a, "a", torch.bfloat16)
    M, K = a.shape
    out = torch.empty(M, n_out, dtype=out_dtype, device=a.device)
    if resid is not None:
        _check(resid, "resid", torch.float32, (M, n_out))
    lib = _lib.load()
    with torch.cuda.device(a.device):
        rc = lib.axvs_linear(a.data_ptr(), K, M, K, w_packed.data_ptr(), _ptr(bias), n_out, scale, int(relu), out.data_ptr(),
                             n_out, int(out_dtype == torch.bfloat16), _ptr(resid), _stream(a.device))
    _lib.check(rc, "axvs_linear")
    return out


def spatial_attention(qkv: torch.Tensor, num_seq: int, F: int, n: int) -> torch.Tensor:
    """Per-frame-softmax attention: qkv bf16 [num_seq*F*n, 768] -> x bf16 [num_seq*F*n, F, 256]."""
    _check(qkv, "qkv", torch.bfloat16, (num_seq * F * n, 768))
    x = torch.empty(num_seq * F * n, F, C, dtype=torch.bfloat16, device=qkv.device)
    lib = _lib.load()
    with torch.cuda.device(qkv.device):
        _lib.check(lib.axvs_spatial_attention(qkv.data_ptr(), x.data_ptr(), num_seq, F, n, _stream(qkv.device)),
                   "axvs_spatial_attention")
    return x


def traj_attn_fwd(q_in: torch.Tensor, k_in: torch.Tensor, v_in: torch.Tensor, pos: Optional[torch.Tensor],
                  resid: Optional[torch.Tensor], w: PackedTA, B: int, T: int, H: int, W: int, axis: int) -> torch.Tensor:
    """out = resid + TrajectoryAttention(q_in + pos, k_in + pos, v_in) on canonical [(B T)(H W), 256] fp32 tensors."""
    rows = B * T * H * W
    for nm, t in (("q_in", q_in), ("k_in", k_in), ("v_in", v_in)):
        _check(t, nm, torch.float32)
        if t.numel() != rows * C:
            raise RuntimeError(f"{nm} has {t.numel()} elements, expected {rows}x{C}")
    pos_clips = 0
    if pos is not None:
        _check(pos, "pos", torch.float32)
        if pos.numel() == rows * C:
            pos_clips = B
        elif pos.numel() == T * H * W * C:
            pos_clips = 1                         # one table shared by all clips
        else:
            raise RuntimeError("pos size mismatch")
    if resid is not None:
        _check(resid, "resid", torch.float32)
    out = torch.empty(rows, C, dtype=torch.float32, device=q_in.device)
    lib = _lib.load()
    nbytes = lib.axvs_traj_attn_workspace_bytes(B, T, H, W)
    with torch.cuda.device(q_in.device):
        ws = workspace(nbytes, q_in.device)
        st = w.struct()
        rc = lib.axvs_traj_attn_fwd(q_in.data_ptr(), k_in.data_ptr(), v_in.data_ptr(), _ptr(pos), pos_clips, _ptr(resid), out.data_ptr(),
                                    ctypes.byref(st), B, T, H, W, axis, ws.data_ptr(), ws.numel(), _stream(q_in.device))
    _lib.check(rc, "axvs_traj_attn_fwd")
    return out


def traj_attn_maps(q_in: torch.Tensor, pos: Optional[torch.Tensor], w: PackedTA, B: int, T: int, H: int, W: int, axis: int) -> torch.Tensor:
    """The reference's `space_attn` maps: fp32 [(num_seq*8), N, F, n] (slow path for the attention visualiser)."""
    rows = B * T * H * W
    _check(q_in, "q_in", torch.float32)
    if q_in.numel() != rows * C:
        raise RuntimeError(f"q_in has {q_in.numel()} elements, expected {rows}x{C}")
    if pos is not None:
        _check(pos, "pos", torch.float32)
        if pos.numel() != rows * C:                   # this slow path reads one positional row per token: materialise a shared table
            if pos.numel() != T * H * W * C:
                raise RuntimeError("pos size mismatch")
            pos = pos.reshape(1, T * H * W, C).expand(B, -1, -1).contiguous()
    num_seq, n = {AXIS_H: (B * W, H), AXIS_W: (B * H, W), AXIS_NONE: (B, H * W)}[axis]
    N = T * n
    maps = torch.empty(num_seq * HEADS, N, T, n, dtype=torch.float32, device=q_in.device)
    lib = _lib.load()
    nbytes = lib.axvs_traj_attn_workspace_bytes(B, T, H, W)
    with torch.cuda.device(q_in.device):
        ws = workspace(nbytes, q_in.device)
        st = w.struct()
        rc = lib.axvs_traj_attn_maps(q_in.data_ptr(), q_in.data_ptr(), _ptr(pos), maps.data_ptr(), ctypes.byref(st), B, T, H, W, axis,
                                     ws.data_ptr(), ws.numel(), _stream(q_in.device))
    _lib.check(rc, "axvs_traj_attn_maps")
    return maps


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    _check(x, "x", torch.float32)
    rows = x.numel() // C
    y = torch.empty_like(x)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.axvs_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), None, rows, eps,
                                      _stream(x.device)), "axvs_layernorm")
    return y


def shared_pos(pos: torch.Tensor) -> torch.Tensor:
    """A positional tensor that is a stride-0 broadcast over the clip dim (`table.expand(B, ...)`) -> its single [1, ...] table."""
    if pos.dim() >= 2 and pos.shape[0] > 1 and pos.stride(0) == 0:
        return pos[:1]
    return pos


def axial_layer_fwd(src: torch.Tensor, pos: torch.Tensor, w: PackedLayer, axial: bool = True) -> torch.Tensor:
    """Temporal(Axial)TrajectoryAttentionLayer.forward: src [(B T), (H W), 256], pos [B, T, H, W, 256] (fp32).

    pos may also be [1, T, H, W, 256] (or a stride-0 `expand` of it, see `shared_pos`): the reference's 3-D sine table does not depend
    on the clip index (WC/pos_embeddings.py:86-130), and the kernels then read the one table instead of B copies."""
    _check(src, "src", torch.float32)
    pos = shared_pos(pos)
    _check(pos, "pos", torch.float32)
    if pos.dim() != 5 or pos.shape[-1] != C:
        raise RuntimeError(f"pos must be [B, T, H, W, {C}], got {tuple(pos.shape)}")
    pos_clips, T, H, W, _ = pos.shape
    if T <= 0 or src.dim() != 3 or src.shape[0] % T:
        raise RuntimeError(f"src must be [(B T), (H W), {C}] with T={T}, got {tuple(src.shape)}")
    B = src.shape[0] // T
    if pos_clips not in (1, B):
        raise RuntimeError(f"pos covers {pos_clips} clips, expected {B} or 1 (shared table)")
    if tuple(src.shape) != (B * T, H * W, C):
        raise RuntimeError(f"src must be [(B T)={B * T}, (H W)={H * W}, {C}], got {tuple(src.shape)}")
    if src.device != pos.device:
        raise RuntimeError("src and pos must be on the same device")
    out = torch.empty_like(src)
    lib = _lib.load()
    nbytes = lib.axvs_layer_workspace_bytes(B, T, H, W, w.d_ffn)
    with torch.cuda.device(src.device):
        ws = workspace(nbytes, src.device)
        st = w.struct()
        rc = lib.axvs_axial_layer_fwd(src.data_ptr(), pos.data_ptr(), pos_clips, out.data_ptr(), ctypes.byref(st), B, T, H, W, int(axial),
                                      ws.data_ptr(), ws.numel(), _stream(src.device))
    _lib.check(rc, "axvs_axial_layer_fwd")
    return out


def pos3d(B: int, T: int, H: int, W: int, level_embed: Optional[torch.Tensor], device) -> torch.Tensor:
    """PositionEmbeddingSine3D(128, normalize=True) table + level embed, channels-last fp32 [B,T,H,W,256]."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("pos3d: CUDA device required (no CPU implementation)")
    if level_embed is not None:
        _check(level_embed, "level_embed", torch.float32, (C,))
    out = torch.empty(B, T, H, W, C, dtype=torch.float32, device=device)
    lib = _lib.load()
    with torch.cuda.device(device):
        _lib.check(lib.axvs_pos3d(out.data_ptr(), _ptr(level_embed), B, T, H, W, _stream(device)), "axvs_pos3d")
    return out


# ------------------------------------------------------------------------------------------------ torch.library
# Thin custom-op registration so the kernels are visible as torch.ops.axialvs.* (schema-checked, fake-tensor aware).
@torch.library.custom_op("axialvs::axial_layer_fwd", mutates_args=(), device_types="cuda")
def _axial_layer_op(src: torch.Tensor, pos: torch.Tensor, weights: Sequence[torch.Tensor], d_ffn: int, axial: bool) -> torch.Tensor:
    return axial_layer_fwd(src, pos, PackedLayer.from_tensors(list(weights), d_ffn), axial)


@_axial_layer_op.register_fake
def _(src, pos, weights, d_ffn, axial):
    return torch.empty_like(src)


@torch.library.custom_op("axialvs::pos3d", mutates_args=(), device_types="cuda")
def _pos3d_op(like: torch.Tensor, level_embed: torch.Tensor, B: int, T: int, H: int, W: int) -> torch.Tensor:
    return pos3d(B, T, H, W, level_embed, like.device)


@_pos3d_op.register_fake
def _(like, level_embed, B, T, H, W):
    return like.new_empty(B, T, H, W, C, dtype=torch.float32)


def ln_ffn_fwd(x: torch.Tensor, w: PackedLayer) -> torch.Tensor:
    """out = LN2(s + W2 relu(W1 s + b1) + b2), s = LN1(x);  x fp32 [rows, 256]."""
    _check(x, "x", torch.float32)
    rows = x.numel() // C
    out = torch.empty_like(x)
    lib = _lib.load()
    nbytes = lib.axvs_ffn_workspace_bytes(rows, w.d_ffn)
    with torch.cuda.device(x.device):
        ws = workspace(nbytes, x.device)
        st = w.struct()
        rc = lib.axvs_ln_ffn_fwd(x.data_ptr(), out.data_ptr(), ctypes.byref(st), rows, ws.data_ptr(), ws.numel(), _stream(x.device))
    _lib.check(rc, "axvs_ln_ffn_fwd")
    return out


DEFAULT_FUSION = 4     # fastest measured level (include/axvs.h lists them); 5 = attention inside the q|k|v kernel, opt-in


def set_fusion(level: int) -> int:
    """Select the fusion level of the composite calls (0 = unfused validation baseline ... 4, the default); returns the previous level."""
    return _lib.load().axvs_set_fusion(int(level))


def set_pair_mode(mask: int) -> int:
    """CTA-pair (cta_group::2) kernels, bit mask: 2 = temporal kernel, 4 = q|k|v projection, 8 = FFN; 16 = frame-major row order between
    the attention and the temporal kernel (no x_diag image); default 30; 0 = single-CTA kernels, pass-order rows.  Returns the previous mask."""
    return _lib.load().axvs_set_pair_mode(int(mask))


def set_attn_core(core: int) -> int:
    """1 = tcgen05 attention core (default), 0 = the mma.sync kernels (validation baseline); returns the previous setting."""
    return _lib.load().axvs_set_attn_core(int(core))


# ------------------------------------------------------------------------------------------------ measurement hooks
def profile_enable(on: bool) -> None:
    """Reset the library's launch counters; with on=True every launch is bracketed by CUDA events (bench.py)."""
    _lib.check(_lib.load().axvs_profile_enable(int(on)), "axvs_profile_enable")


def profile_read() -> Dict[str, Dict[str, float]]:
    """Per kernel class: device ms, algorithmic flops/bytes, launches since enable, event-timed launches."""
    lib = _lib.load()
    n = lib.axvs_profile_num_classes()
    ms, fl, by = (ctypes.c_double * n)(), (ctypes.c_double * n)(), (ctypes.c_double * n)()
    la, ti = (ctypes.c_longlong * n)(), (ctypes.c_longlong * n)()
    _lib.check(lib.axvs_profile_read(ms, fl, by, la, ti), "axvs_profile_read")
    return {lib.axvs_profile_class_name(i).decode(): dict(ms=ms[i], flops=fl[i], bytes=by[i], launches=int(la[i]), timed=int(ti[i]))
            for i in range(n)}


# ------------------------------------------------------------------------------------------------ cross-clip tail
def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 [rows, 256] -> bf16 [rows, 256]."""
    _check(x, "x", torch.float32)
    rows = x.numel() // C
    out = torch.empty(rows, C, dtype=torch.bfloat16, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.axvs_cast_bf16(x.data_ptr(), out.data_ptr(), rows, _stream(x.device)), "axvs_cast_bf16")
    return out


@dataclass
class PackedAspp:
    """struct axvs_aspp_weights (+ the tensors that keep the pointers alive)."""
    w_conv: List[torch.Tensor]
    b_conv: List[torch.Tensor]
    dilation: List[int]
    w_proj: torch.Tensor
    lncf_g: torch.Tensor
    lncf_b: torch.Tensor
    ln_g: torch.Tensor
    ln_b: torch.Tensor
    split: bool = False

    def struct(self) -> "_lib.AsppWeights":
        s = _lib.AsppWeights()
        for i in range(3):
            s.w_conv[i] = self.w_conv[i].data_ptr()
            s.b_conv[i] = self.b_conv[i].data_ptr()
            s.dilation[i] = int(self.dilation[i])
        s.w_proj = self.w_proj.data_ptr()
        s.lncf_g, s.lncf_b, s.ln_g, s.ln_b = (t.data_ptr() for t in (self.lncf_g, self.lncf_b, self.ln_g, self.ln_b))
        s.split = int(self.split)
        return s


def pack_weight_split(w: torch.Tensor) -> torch.Tensor:
    """Split-precision image of an fp32 weight [n_out, k]: [W | W | W - bf16(W)] packed as [n_out, 3k] (include/axvs.h, axvs_linear_f32)."""
    w = w.detach().float()
    return pack_weight(torch.cat((w, w, w - w.bfloat16().float()), dim=1).contiguous())


def pack_aspp(p: Dict[str, torch.Tensor], ln_w: torch.Tensor, ln_b: torch.Tensor, atrous_rates: Sequence[int], split: bool = True) -> PackedAspp:
    """Pack an `ASPP` state dict (CC:176-201) + the following `conv_norms[i]` LayerNorm.  split=True (default): split-precision
    weight images, fp32-grade GEMMs (the rows are clips x queries: the 3x tensor work is negligible)."""
    def g(name):
        return p[name].detach().float().contiguous()

    pw = pack_weight_split if split else pack_weight

    wc, bc = [], []
    for i in range(3):
        w = g(f"_aspp_conv{i}.weight")                                  # [256, 256, 3] (out, in, tap)
        if w.shape[-1] != 3:
            raise NotImplementedError("axial_vs_b200: ASPP kernel_size must be 3 (every shipped config)")
        wc.append(pw(w.permute(0, 2, 1).reshape(256, 768).contiguous()))   # K index = tap * 256 + c_in
        bc.append(g(f"_aspp_conv{i}.bias"))
    wproj = pw(g("_proj_conv_bn_act.conv.weight")[:, :, 0].contiguous())   # [256, 768]
    return PackedAspp(wc, bc, [int(r) for r in atrous_rates], wproj, g("_proj_conv_bn_act.norm.weight"), g("_proj_conv_bn_act.norm.bias"),
                      ln_w.detach().float().contiguous(), ln_b.detach().float().contiguous(), bool(split))


def cc_aspp_fwd(x: torch.Tensor, w: PackedAspp, b: int, T: int, Q: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """out = LN(GELU(LN_cf(proj(cat_d conv_d(z)))) + z) on rows (b, t, q); returns (fp32, bf16)."""
    _check(x, "x", torch.float32)
    rows = b * T * Q
    if x.numel() != rows * C:
        raise RuntimeError("cc_aspp_fwd: x size mismatch")
    out = torch.empty(rows, C, dtype=torch.float32, device=x.device)
    out16 = torch.empty(rows, C, dtype=torch.bfloat16, device=x.device)
    lib = _lib.load()
    nbytes = lib.axvs_cc_aspp_workspace_bytes(rows)
    with torch.cuda.device(x.device):
        ws = workspace(nbytes, x.device)
        st = w.struct()
        rc = lib.axvs_cc_aspp_fwd(x.data_ptr(), out.data_ptr(), out16.data_ptr(), ctypes.byref(st), b, T, Q, ws.data_ptr(), ws.numel(),
                                  _stream(x.device))
    _lib.check(rc, "axvs_cc_aspp_fwd")
    return out, out16


def linear_act(a: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], n_out: int, act: int,
               out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """axvs_linear with activation code (0 none, 1 ReLU, 2 GELU)."""
    _check(a, "a", torch.bfloat16)
    M, K = a.shape
    out = torch.empty(M, n_out, dtype=out_dtype, device=a.device)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        rc = lib.axvs_linear(a.data_ptr(), K, M, K, w_packed.data_ptr(), _ptr(bias), n_out, 1.0, int(act), out.data_ptr(), n_out,
                             int(out_dtype == torch.bfloat16), None, _stream(a.device))
    _lib.check(rc, "axvs_linear")
    return out


def linear_f32(a: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], n_out: int, act: int = 0, split: bool = True,
               out_dtype: torch.dtype = torch.float32, scale: float = 1.0) -> torch.Tensor:
    """act((a W^T + bias) * scale) on fp32 rows a [M, K]; split=True: w_packed from `pack_weight_split`, fp32-grade product (axvs_linear_f32)."""
    _check(a, "a", torch.float32)
    M, K = a.shape
    out = torch.empty(M, n_out, dtype=out_dtype, device=a.device)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        rc = lib.axvs_linear_f32(a.data_ptr(), K, M, K, w_packed.data_ptr(), int(split), _ptr(bias), n_out, float(scale), int(act), out.data_ptr(), n_out,
                                 int(out_dtype == torch.bfloat16), _stream(a.device))
    _lib.check(rc, "axvs_linear_f32")
    return out


def cc_class_pool(ce: torch.Tensor, w_act: torch.Tensor, b_act: float, T: int, Q: int) -> torch.Tensor:
    _check(ce, "ce", torch.bfloat16, (T * Q, C))
    out = torch.empty(Q, C, dtype=torch.bfloat16, device=ce.device)
    lib = _lib.load()
    with torch.cuda.device(ce.device):
        _lib.check(lib.axvs_cc_class_pool(ce.data_ptr(), w_act.data_ptr(), float(b_act), out.data_ptr(), T, Q, _stream(ce.device)),
                   "axvs_cc_class_pool")
    return out


def _level_slice_stride(t: torch.Tensor, name: str, images: int, hw: int) -> int:
    """t fp32 [images, hw, 256], dense, or one level's slice of a multi-level token tensor [images, len, 256] -> floats between images."""
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and tuple(t.shape) == (images, hw, C)):
        raise RuntimeError(f"{name} must be a CUDA fp32 tensor of shape {(images, hw, C)}")
    if t.stride(2) != 1 or t.stride(1) != C or (images > 1 and (t.stride(0) < hw * C or t.stride(0) % 4)) or t.data_ptr() % 16:
        raise RuntimeError(f"{name} must be dense or a level slice [:, a:b] of a contiguous [images, len, 256] tensor")
    return t.stride(0) if images > 1 else hw * C


def input_proj_fwd(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, gn_w: torch.Tensor, gn_b: torch.Tensor, eps: float = 1e-5,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x fp32 NCHW [images, c_in, H, W] -> GroupNorm(32)(Conv1x1(x)) as token-major fp32 [images, H*W, 256]; `out`: write into this
    tensor instead (dense, or the level's slice of the multi-level token tensor [images, len, 256])."""
    _check(x, "x", torch.float32)
    if x.dim() != 4:
        raise RuntimeError("input_proj_fwd: expected an NCHW feature map")
    images, c_in, H, W = x.shape
    for t, nm in ((bias, "bias"), (gn_w, "GroupNorm weight"), (gn_b, "GroupNorm bias")):
        _check(t, nm, torch.float32, (C,))
    if out is None:
        out = torch.empty(images, H * W, C, dtype=torch.float32, device=x.device)
    stride = _level_slice_stride(out, "out", images, H * W)
    lib = _lib.load()
    nbytes = lib.axvs_proj_workspace_bytes(images)
    with torch.cuda.device(x.device):
        ws = workspace(nbytes, x.device)
        rc = lib.axvs_input_proj_fwd(x.data_ptr(), w_packed.data_ptr(), bias.data_ptr(), gn_w.data_ptr(), gn_b.data_ptr(), out.data_ptr(), stride,
                                     images, c_in, H * W, float(eps), ws.data_ptr(), ws.numel(), _stream(x.device))
    _lib.check(rc, "axvs_input_proj_fwd")
    return out


def output_proj_fwd(tokens: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, gn_w: torch.Tensor, gn_b: torch.Tensor,
                    H: int, W: int, eps: float = 1e-5) -> torch.Tensor:
    """tokens fp32 [images, H*W, 256] (dense, or a level slice of the multi-level token tensor) -> GroupNorm(32)(Conv1x1(tokens^T)) as fp32
    NCHW [images, c_out, H, W]."""
    if tokens.dim() != 3 or tokens.shape[1] != H * W or tokens.shape[2] != C:
        raise RuntimeError("output_proj_fwd: expected tokens [images, H*W, 256]")
    images, c_out = tokens.shape[0], gn_w.numel()
    stride = _level_slice_stride(tokens, "tokens", images, H * W)
    _check(bias, "bias", torch.float32, ((c_out + 255) // 256 * 256,))          # zero-padded to the GEMM's 256-column chunks
    for t, nm in ((gn_w, "GroupNorm weight"), (gn_b, "GroupNorm bias")):
        _check(t, nm, torch.float32, (c_out,))
    out = torch.empty(images, c_out, H, W, dtype=torch.float32, device=tokens.device)
    lib = _lib.load()
    with torch.cuda.device(tokens.device):
        rc = lib.axvs_output_proj_fwd(tokens.data_ptr(), stride, w_packed.data_ptr(), bias.data_ptr(), gn_w.data_ptr(), gn_b.data_ptr(), out.data_ptr(),
                                      images, c_out, H * W, float(eps), _stream(tokens.device))
    _lib.check(rc, "axvs_output_proj_fwd")
    return out


@dataclass
class PackedMsda:
    """struct axvs_msda_weights (+ the tensors that keep the pointers alive)."""
    tensors: List[torch.Tensor]
    d_ffn: int
    n_levels: int
    n_points: int

    def struct(self) -> "_lib.MsdaWeights":
        return _lib.MsdaWeights(*[t.data_ptr() for t in self.tensors], self.d_ffn, self.n_levels, self.n_points)


def pack_msda_layer(p: Dict[str, torch.Tensor], n_levels: int, n_points: int) -> PackedMsda:
    """Pack the state dict of MSDeformAttnTransformerEncoderLayer (WC/msdeformattn.py:177-203)."""
    def g(name):
        return p[name].detach().float().contiguous()

    no = HEADS * n_levels * n_points
    w_oa = torch.zeros(512, C, device=g("self_attn.value_proj.weight").device)
    b_oa = torch.zeros(512, device=w_oa.device)
    w_oa[: 2 * no] = g("self_attn.sampling_offsets.weight")
    w_oa[2 * no: 3 * no] = g("self_attn.attention_weights.weight")
    b_oa[: 2 * no] = g("self_attn.sampling_offsets.bias")
    b_oa[2 * no: 3 * no] = g("self_attn.attention_weights.bias")
    ts = [pack_weight(g("self_attn.value_proj.weight")), g("self_attn.value_proj.bias"), pack_weight(w_oa), b_oa,
          pack_weight(g("self_attn.output_proj.weight")), g("self_attn.output_proj.bias"), g("norm1.weight"), g("norm1.bias"),
          pack_weight(g("linear1.weight")), g("linear1.bias"), pack_weight(g("linear2.weight")), g("linear2.bias"),
          pack_weight_units(g("linear1.weight")), pack_weight_units(g("linear2.weight"), k_major=1),
          pack_weight_units(g("linear1.weight"), k_major=2), g("norm2.weight"), g("norm2.bias")]
    # fused front end (msda_front_pair_kernel): [offsets | logits rows zero-padded to 384 ; value_proj] as one unit image + bias
    w_front = torch.cat((w_oa[:384], g("self_attn.value_proj.weight")), 0).contiguous()
    b_front = torch.cat((b_oa[:384], g("self_attn.value_proj.bias"))).contiguous()
    ts += [pack_weight_units(w_front), b_front, pack_weight_units(g("self_attn.output_proj.weight"))]
    return PackedMsda(ts, p["linear1.weight"].shape[0], n_levels, n_points)


def pack_msda_sampler(sampling_offsets_w, sampling_offsets_b, attention_weights_w, attention_weights_b, value_w, value_b, n_levels: int, n_points: int):
    """(w_value, b_value, w_oa, b_oa) images for `msda_sample_fwd` from the four Linear layers of a deformable attention."""
    no = HEADS * n_levels * n_points
    w_oa = torch.zeros(512, C, device=value_w.device)
    b_oa = torch.zeros(512, device=value_w.device)
    w_oa[: 2 * no] = sampling_offsets_w.detach().float()
    w_oa[2 * no: 3 * no] = attention_weights_w.detach().float()
    b_oa[: 2 * no] = sampling_offsets_b.detach().float()
    b_oa[2 * no: 3 * no] = attention_weights_b.detach().float()
    return pack_weight(value_w.detach().float().contiguous()), value_b.detach().float().contiguous(), pack_weight(w_oa), b_oa


def msda_sample_fwd(value_in: torch.Tensor, query_in: torch.Tensor, pos: Optional[torch.Tensor], ref_points: torch.Tensor,
                    shapes: Sequence[Tuple[int, int]], packed, n_levels: int, n_points: int) -> torch.Tensor:
    """Sampling half of a deformable attention: value_in / query_in fp32 [images, len, 256], pos [images or 1, len, 256] or None,
    ref_points [images or 1, len, n_levels, 2] -> sampled bf16 [images, len, 256]."""
    _check(value_in, "value", torch.float32)
    _check(query_in, "query", torch.float32)
    images, length, c = value_in.shape
    if c != C or tuple(query_in.shape) != (images, length, C):
        raise RuntimeError("msda_sample_fwd: value and query must be [images, len, 256]")
    if pos is not None:
        _check(pos, "pos", torch.float32)
        if pos.dim() != 3 or pos.shape[0] not in (1, images) or tuple(pos.shape[1:]) != (length, C):
            raise RuntimeError("msda_sample_fwd: pos must be [images or 1, len, 256]")
    _check(ref_points, "reference_points", torch.float32)
    if ref_points.dim() != 4 or ref_points.shape[0] not in (1, images) or tuple(ref_points.shape[1:]) != (length, n_levels, 2):
        raise RuntimeError(f"msda_sample_fwd: reference_points must be [{images} or 1, {length}, {n_levels}, 2]")
    if len(shapes) != n_levels or sum(int(h) * int(v) for h, v in shapes) != length:
        raise RuntimeError("msda_sample_fwd: spatial shapes do not match the token count")
    out = torch.empty(images, length, C, dtype=torch.bfloat16, device=value_in.device)
    lib = _lib.load()
    hw = (ctypes.c_int * (2 * n_levels))(*[int(v) for s_ in shapes for v in s_])
    wv, bv, woa, boa = packed
    nbytes = lib.axvs_msda_sample_workspace_bytes(images * length)
    with torch.cuda.device(value_in.device):
        ws = workspace(nbytes, value_in.device)
        rc = lib.axvs_msda_sample_fwd(value_in.data_ptr(), query_in.data_ptr(), _ptr(pos), pos.shape[0] if pos is not None else 0, ref_points.data_ptr(),
                                      ref_points.shape[0], hw, wv.data_ptr(), bv.data_ptr(), woa.data_ptr(), boa.data_ptr(), n_levels, n_points,
                                      out.data_ptr(), images, length, ws.data_ptr(), ws.numel(), _stream(value_in.device))
    _lib.check(rc, "axvs_msda_sample_fwd")
    return out


def msda_layer_fwd(src: torch.Tensor, pos: Optional[torch.Tensor], ref_points: torch.Tensor, shapes: Sequence[Tuple[int, int]],
                   w: PackedMsda) -> torch.Tensor:
    """MSDeformAttn spatial encoder layer: src fp32 [images, len, 256]; pos fp32 [images or 1, len, 256]; ref_points fp32
    [images or 1, len, n_levels, 2] (a leading dimension of 1 is broadcast over the images inside the kernels)."""
    _check(src, "src", torch.float32)
    images, length, c = src.shape
    if c != C:
        raise RuntimeError("msda_layer_fwd: 256 channels expected")
    if pos is not None:
        _check(pos, "pos", torch.float32)
        if pos.dim() != 3 or pos.shape[0] not in (1, images) or tuple(pos.shape[1:]) != (length, C):
            raise RuntimeError(f"msda_layer_fwd: pos must be [{images} or 1, {length}, {C}], got {tuple(pos.shape)}")
    _check(ref_points, "reference_points", torch.float32)
    if ref_points.dim() != 4 or ref_points.shape[0] not in (1, images) or tuple(ref_points.shape[1:]) != (length, w.n_levels, 2):
        raise RuntimeError(f"msda_layer_fwd: reference_points must be [{images} or 1, {length}, {w.n_levels}, 2]")
    if len(shapes) != w.n_levels or sum(int(h) * int(v) for h, v in shapes) != length:
        raise RuntimeError("msda_layer_fwd: spatial shapes do not match the token count")
    out = torch.empty_like(src)
    lib = _lib.load()
    hw = (ctypes.c_int * (2 * w.n_levels))(*[int(v) for s_ in shapes for v in s_])
    nbytes = lib.axvs_msda_layer_workspace_bytes(images * length, w.d_ffn)
    with torch.cuda.device(src.device):
        ws = workspace(nbytes, src.device)
        st = w.struct()
        rc = lib.axvs_msda_layer_fwd(src.data_ptr(), pos.data_ptr() if pos is not None else None, pos.shape[0] if pos is not None else 0,
                                     ref_points.data_ptr(), ref_points.shape[0], hw, out.data_ptr(), ctypes.byref(st), images, length,
                                     ws.data_ptr(), ws.numel(), _stream(src.device))
    _lib.check(rc, "axvs_msda_layer_fwd")
    return out


def query_self_attn(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, sim_affine: torch.Tensor, val_affine: torch.Tensor) -> torch.Tensor:
    """AttentionOperation core: q, k fp32 [N, heads, 16, L], v fp32 [N, heads, 32, L] -> GELU(BN(softmax(BN(q.k)) v)) fp32 [N, heads*32, L]."""
    for t, nm in ((q, "query"), (k, "key"), (v, "value"), (sim_affine, "sim_affine"), (val_affine, "val_affine")):
        _check(t, nm, torch.float32)
    if q.dim() != 4 or q.shape != k.shape or v.dim() != 4 or v.shape[:2] != q.shape[:2] or v.shape[3] != q.shape[3]:
        raise RuntimeError("query_self_attn: expected query/key [N, heads, dk, L] and value [N, heads, dv, L]")
    N, heads, dk, L = q.shape
    if dk != 16 or v.shape[2] != 32:
        raise RuntimeError(f"query_self_attn: built for key depth 16 and value depth 32 per head (got {dk}, {v.shape[2]})")
    if sim_affine.numel() != 2 * heads or val_affine.numel() != 2 * heads * 32:
        raise RuntimeError("query_self_attn: folded BatchNorm sizes do not match the number of heads")
    out = torch.empty(N, heads * 32, L, dtype=torch.float32, device=q.device)
    lib = _lib.load()
    with torch.cuda.device(q.device):
        rc = lib.axvs_query_self_attn(q.data_ptr(), k.data_ptr(), v.data_ptr(), sim_affine.data_ptr(), val_affine.data_ptr(), out.data_ptr(),
                                      N, heads, L, _stream(q.device))
    _lib.check(rc, "axvs_query_self_attn")
    return out


def cm_to_rows(x: torch.Tensor, gelu: bool = False) -> torch.Tensor:
    """fp32 [N, C, M] (channel-major, the reference's layout) -> token rows [N*M, C]; optional GELU on the way (axvs_cm_to_rows)."""
    _check(x, "x", torch.float32)
    N, Cc, M = x.shape
    rows = torch.empty(N * M, Cc, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.axvs_cm_to_rows(x.data_ptr(), rows.data_ptr(), N, Cc, M, 2 if gelu else 0, _stream(x.device)), "axvs_cm_to_rows")
    return rows


def rows_to_cm(rows: torch.Tensor, N: int, channels: int, normalize: bool = False) -> torch.Tensor:
    """token rows fp32 [N*M, ld] -> [N, channels, M] (first `channels` columns); normalize: L2 over the channels (axvs_rows_to_cm)."""
    _check(rows, "rows", torch.float32)
    M = rows.shape[0] // N
    out = torch.empty(N, channels, M, dtype=torch.float32, device=rows.device)
    lib = _lib.load()
    with torch.cuda.device(rows.device):
        _lib.check(lib.axvs_rows_to_cm(rows.data_ptr(), rows.shape[1], out.data_ptr(), N, channels, M, int(normalize), _stream(rows.device)), "axvs_rows_to_cm")
    return out


def dwconv5(x_rows: torch.Tensor, w: torch.Tensor, affine: torch.Tensor, N: int, H: int, W: int, act: int = 2) -> torch.Tensor:
    """Depthwise 5x5 convolution (padding 2) + folded batch norm + activation on channels-last rows [N*H*W, C] (axvs_dwconv5)."""
    for t, nm in ((x_rows, "x"), (w, "w"), (affine, "affine")):
        _check(t, nm, torch.float32)
    Cc = x_rows.shape[1]
    if x_rows.shape[0] != N * H * W or w.numel() != Cc * 25 or affine.numel() != 2 * Cc:
        raise RuntimeError("dwconv5: size mismatch")
    y = torch.empty_like(x_rows)
    lib = _lib.load()
    with torch.cuda.device(x_rows.device):
        _lib.check(lib.axvs_dwconv5(x_rows.data_ptr(), w.data_ptr(), affine.data_ptr(), y.data_ptr(), N, H, W, Cc, act, _stream(x_rows.device)), "axvs_dwconv5")
    return y


def add_act(a: torch.Tensor, b: Optional[torch.Tensor], act: int = 0) -> torch.Tensor:
    """act(a + b) elementwise on fp32 tensors of equal shape (axvs_add_act); act 0 none, 1 ReLU, 2 GELU."""
    _check(a, "a", torch.float32)
    if b is not None:
        _check(b, "b", torch.float32, tuple(a.shape))
    y = torch.empty_like(a)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.axvs_add_act(a.data_ptr(), _ptr(b), y.data_ptr(), a.numel(), act, _stream(a.device)), "axvs_add_act")
    return y


def masked_mha(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: Optional[torch.Tensor], heads: int = 8, seq_first: bool = True,
               out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """softmax(q k^T + mask) v per head (axvs_masked_mha_fwd).  q [Nq, B, C] / k, v [L, B, C] fp32 (seq_first) or [B, N, C]; q already carries
    head_dim^-0.5 * log2(e); mask bool / uint8 [B*heads, Nq, L] (True = blocked) or None.  Returns the heads' outputs in q's layout."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        _check(t, nm, torch.float32)
    if q.dim() != 3 or k.shape != v.shape or k.dim() != 3 or q.shape[2] != k.shape[2] or q.shape[2] != heads * 32:
        raise RuntimeError("masked_mha: expected q [Nq, B, heads*32] and k, v [L, B, heads*32] (or batch-first)")
    (Nq, B), L = (q.shape[:2], k.shape[0]) if seq_first else ((q.shape[1], q.shape[0]), k.shape[1])
    if (k.shape[1] if seq_first else k.shape[0]) != B:
        raise RuntimeError("masked_mha: batch sizes of q and k differ")
    mptr = None
    if mask is not None:
        if mask.dtype == torch.bool:
            mask = mask.view(torch.uint8)
        _check(mask, "mask", torch.uint8, (B * heads, Nq, L))
        mptr = mask.data_ptr()
    out = torch.empty_like(q, dtype=out_dtype)
    lib = _lib.load()
    nbytes = lib.axvs_masked_mha_workspace_bytes(B, heads, Nq, L)
    ws = workspace(nbytes, q.device)
    with torch.cuda.device(q.device):
        rc = lib.axvs_masked_mha_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), mptr, out.data_ptr() if out_dtype == torch.float32 else None,
                                     out.data_ptr() if out_dtype == torch.bfloat16 else None, B, heads, Nq, L, int(seq_first), ws.data_ptr(), nbytes,
                                     _stream(q.device))
    _lib.check(rc, "axvs_masked_mha_fwd")
    return out


def frame_attn_f32(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, F: int, heads: int = 8) -> torch.Tensor:
    """Per-frame-softmax attention in fp32: q (pre-scaled by head_dim^-0.5 * log2 e), k, v [B, N, heads*32], N = F * n -> x [B, N, F, heads*32]."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        _check(t, nm, torch.float32, tuple(q.shape))
    B, N, Cc = q.shape
    if Cc != heads * 32 or N % F:
        raise RuntimeError("frame_attn_f32: expected [B, F*n, heads*32]")
    x = torch.empty(B, N, F, Cc, dtype=torch.float32, device=q.device)
    lib = _lib.load()
    nbytes = lib.axvs_frame_attn_f32_workspace_bytes(B, heads, N, F)
    ws = workspace(nbytes, q.device)
    with torch.cuda.device(q.device):
        _lib.check(lib.axvs_frame_attn_f32(q.data_ptr(), k.data_ptr(), v.data_ptr(), x.data_ptr(), B, heads, N, F, N // F, ws.data_ptr(), nbytes,
                                           _stream(q.device)), "axvs_frame_attn_f32")
    return x


def kmeans_update(mask_logits: torch.Tensor, pixel_value: torch.Tensor, advanced: bool = False,
                  return_assignment: bool = False):
    """k-means cluster update: mask_logits fp32 [N, L, M], pixel_value fp32 [N, 256, M] -> fp32 [N, 256, L] (and int32 [N, M])."""
    _check(mask_logits, "mask_logits", torch.float32)
    _check(pixel_value, "pixel_value", torch.float32)
    if mask_logits.dim() != 3 or pixel_value.dim() != 3 or pixel_value.shape[0] != mask_logits.shape[0] or \
            pixel_value.shape[2] != mask_logits.shape[2] or pixel_value.shape[1] != C:
        raise RuntimeError("kmeans_update: expected mask_logits [N, L, M] and pixel_value [N, 256, M]")
    N, L, M = mask_logits.shape
    out = torch.empty(N, C, L, dtype=torch.float32, device=mask_logits.device)
    assign = torch.empty(N, M, dtype=torch.int32, device=mask_logits.device) if return_assignment else None
    lib = _lib.load()
    nbytes = lib.axvs_kmeans_update_workspace_bytes(N, L, M)
    with torch.cuda.device(mask_logits.device):
        ws = workspace(nbytes, mask_logits.device)
        rc = lib.axvs_kmeans_update(mask_logits.data_ptr(), pixel_value.data_ptr(), out.data_ptr(),
                                    assign.data_ptr() if assign is not None else None, N, L, M, int(bool(advanced)),
                                    ws.data_ptr(), ws.numel(), _stream(mask_logits.device))
    _lib.check(rc, "axvs_kmeans_update")
    return (out, assign) if return_assignment else out


def mask_einsum(pixel: torch.Tensor, mk: torch.Tensor, T: int, Q: int, P: int, bn_scale: float, bn_shift: float,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[q, t, p] = bn_scale * sum_c pixel[t, c, p] mk[t*Q + q, c] + bn_shift; pixel fp32 [T, 128, P], mk [T*Q, ld >= 128]:
    bf16 (plain bf16 products) or fp32 (split-precision products, fp32-grade logits)."""
    _check(pixel, "pixel", torch.float32)
    if mk.dtype not in (torch.bfloat16, torch.float32):
        raise RuntimeError("mask_einsum: mk must be bf16 or fp32")
    _check(mk, "mk", mk.dtype)
    channels = pixel.numel() // (T * P)
    if pixel.numel() != T * channels * P or mk.shape[0] != T * Q or (mk.dtype == torch.bfloat16 and channels != 128):
        raise RuntimeError("mask_einsum: size mismatch")
    if out is None:
        out = torch.empty(Q, T, P, dtype=torch.float32, device=pixel.device)
    else:
        _check(out, "out", torch.float32)
        if out.numel() != Q * T * P:
            raise RuntimeError("mask_einsum: out has the wrong size")
    lib = _lib.load()
    with torch.cuda.device(pixel.device):
        if mk.dtype == torch.float32:
            rc = lib.axvs_mask_einsum_f32(pixel.data_ptr(), mk.data_ptr(), mk.shape[1], out.data_ptr(), T, Q, P, channels, float(bn_scale),
                                          float(bn_shift), _stream(pixel.device))
        else:
            rc = lib.axvs_mask_einsum(pixel.data_ptr(), mk.data_ptr(), mk.shape[1], out.data_ptr(), T, Q, P, float(bn_scale), float(bn_shift),
                                      _stream(pixel.device))
    _lib.check(rc, "axvs_mask_einsum")
    return out
