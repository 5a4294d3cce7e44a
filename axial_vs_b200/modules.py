"""Drop-in nn.Modules for the within-clip tracking module's temporal layers (Video-kMaX flavour).

Same class names, constructor arguments, forward signatures, return tuples and state-dict keys as
`MaXTron_Video-kMaX/maxtron_deeplab/modeling/within_clip_tracking_module/temporal_attention.py`
(TrajectoryAttention :20-76, TemporalEncoder :79-100, TemporalTrajectoryAttentionLayer :103-155,
TemporalAxialTrajectoryAttentionLayer :158-220), so `msdeformattn.py:54` can construct them unchanged and
`DetectionCheckpointer` loads the reference checkpoints (keys `...temporal_layers.{i}.height_attn.q.weight` etc.;
the LR-multiplier substring rules on `temporal_layers`, Vk/train_net_video.py:156-163, keep matching).

The forward pass runs entirely in libaxvs.so (sm_100a).  Inference only: the reference's hot path is the
eval forward; calling these modules in training mode with autograd enabled raises.

One intentional deviation ("Attention-map return value"): the `[(B' h), N, F, n]` softmax maps,
consumed only by the attention visualiser (Vk/maxtron_deeplab/maxtron_wc_model.py:598-611), are returned
as `None` unless `module.return_attn_maps = True`.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import ops


def _require_inference(mod: nn.Module, *tensors: Tensor) -> None:
    if torch.is_grad_enabled() and (mod.training or any(t.requires_grad for t in tensors)):
        raise RuntimeError(
            f"{type(mod).__name__}: axial_vs_b200 implements the inference forward only; call .eval() and run under "
            "torch.no_grad() (the reference's training path stays on stock PyTorch)")


class _PackedCache:
    """Packed bf16 weights derived from the fp32 nn.Parameters AND buffers (BatchNorm running statistics are folded into some
    images); rebuilt when any of them is replaced or modified in place through autograd-visible ops (`_version`).  Writes that
    bypass the version counter (`p.data.fill_()`, which the reference's init code uses) need an explicit `invalidate()` --
    `invalidate_packed(module)` does that for a whole module tree and is registered as a load_state_dict post-hook by the drop-ins.

    The pack kernels run on the stream that is current when the cache is (re)built; a CUDA event recorded behind them makes any OTHER
    stream that later fetches the images wait for the build (the per-level / per-chunk side streams of within_clip.py fetch them
    while the first level may still be packing)."""

    def __init__(self):
        self._key = None
        self._val = None
        self._event = None
        self._stream = None

    def invalidate(self) -> None:
        self._key = None

    def get(self, module: nn.Module, device: torch.device, build):
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in module.parameters()) + \
            tuple((b.data_ptr(), b._version) for b in module.buffers())
        dev = torch.device(device)
        if key != self._key:
            self._val = build()
            self._key = key
            if dev.type == "cuda":
                self._stream = torch.cuda.current_stream(dev)
                self._event = torch.cuda.Event()
                self._event.record(self._stream)
        elif self._event is not None:
            cur = torch.cuda.current_stream(dev)
            if torch.cuda.is_current_stream_capturing():
                pass      # no event calls inside a CUDA-graph capture (cudaEventQuery invalidates it); `torch.cuda.graph` synchronises the
                          # device when the capture begins, so a build that preceded the capture has completed
            elif cur != self._stream and not self._event.query():
                cur.wait_event(self._event)
            elif self._event.query():
                self._event = None                             # build finished: nothing to order against any more
        return self._val


def _invalidate_hook(module: nn.Module, incompatible_keys=None) -> None:
    invalidate_packed(module)


def invalidate_packed(module: nn.Module) -> None:
    """Drop every packed-weight image under `module` (call after writing parameters through `.data`)."""
    for m in module.modules():
        for v in vars(m).values():
            if isinstance(v, _PackedCache):
                v.invalidate()
            elif isinstance(v, (list, tuple)):
                for c in v:
                    if isinstance(c, _PackedCache):
                        c.invalidate()


class TrajectoryAttention(nn.Module):
    """WC/temporal_attention.py:20-76.  forward(query, key, value, num_frames) -> (x, space_attn)."""

    def __init__(self, dim, num_heads=8, attn_drop=0.):
        super().__init__()
        if dim != ops.C or num_heads != ops.HEADS:
            raise NotImplementedError(f"axial_vs_b200 kernels are specialised for dim=256, num_heads=8 (got {dim}, {num_heads})")
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.q = nn.Linear(dim, dim, bias=True)
        self.k = nn.Linear(dim, dim, bias=True)
        self.v = nn.Linear(dim, dim, bias=True)
        self.proj_q = nn.Linear(dim, dim, bias=True)
        self.proj_kv = nn.Linear(dim, dim * 2, bias=True)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.return_attn_maps = False
        self._cache = _PackedCache()
        self.register_load_state_dict_post_hook(_invalidate_hook)

    def packed(self, device) -> ops.PackedTA:
        return self._cache.get(self, device, lambda: ops.pack_ta({k: v for k, v in self.state_dict().items()}))

    def forward(self, query, key, value, num_frames=2):
        _require_inference(self, query, key, value)
        Bp, N, C = query.shape
        F = num_frames
        if N % F:
            raise RuntimeError(f"sequence length {N} is not a multiple of num_frames {F}")
        n = N // F
        if query.numel() == 0:                    # empty batch of sequences
            return query.new_empty(Bp, N, C), None
        q = query.contiguous().float()
        k = q if key is query else key.contiguous().float()
        v = q if value is query else value.contiguous().float()
        # sequences are already in storage order: B' "clips" of F frames of n tokens
        out = ops.traj_attn_fwd(q, k, v, None, None, self.packed(query.device), Bp, F, n, 1, ops.AXIS_NONE)
        maps = None
        if self.return_attn_maps:
            if key is not query:
                raise NotImplementedError("return_attn_maps requires key is query (all reference call sites)")
            maps = ops.traj_attn_maps(q, None, self.packed(query.device), Bp, F, n, 1, ops.AXIS_NONE)
        return out.view(Bp, N, C).to(query.dtype), maps


class _LayerBase(nn.Module):
    axial = True

    def _init_common(self, d_model, d_ffn, dropout, attn_drop, activation):
        if activation != "relu":
            raise NotImplementedError("axial_vs_b200: only activation='relu' is implemented (every shipped config uses it)")
        if d_ffn % 256:
            raise NotImplementedError("axial_vs_b200: d_ffn must be a multiple of 256")
        self.dropout1 = nn.Dropout(attn_drop)   # sic: the reference swaps the two rates (WC/temporal_attention.py:164-166)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = torch.nn.functional.relu
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self._cache = _PackedCache()
        self.register_load_state_dict_post_hook(_invalidate_hook)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def packed(self, device) -> ops.PackedLayer:
        return self._cache.get(self, device, lambda: ops.pack_layer(dict(self.state_dict()), axial=self.axial))

    def _run(self, src: Tensor, pos: Tensor) -> Tensor:
        _require_inference(self, src, pos)
        if src.numel() == 0:                      # empty batch of clips: nothing to launch (the reference returns an empty tensor too)
            return src.clone()
        out = ops.axial_layer_fwd(src.contiguous().float(), ops.shared_pos(pos).contiguous().float(), self.packed(src.device), self.axial)
        return out.to(src.dtype)


class TemporalTrajectoryAttentionLayer(_LayerBase):
    """Non-axial layer, WC/temporal_attention.py:103-155.  forward(src, pos) -> (src, None, None)."""
    axial = False

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.0, attn_drop=0.0, activation="relu", n_heads=8):
        super().__init__()
        self.temporal_attn = TrajectoryAttention(d_model, n_heads, dropout)
        self._init_common(d_model, d_ffn, dropout, attn_drop, activation)

    def forward(self, src: Tensor, pos: Tensor):
        return self._run(src, pos), None, None


class TemporalAxialTrajectoryAttentionLayer(_LayerBase):
    """Axial layer, WC/temporal_attention.py:158-220.  forward(src [BT,HW,C], pos [B,T,H,W,C]) -> (src, h_map, w_map)."""
    axial = True

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.0, attn_drop=0.0, activation="relu", n_heads=8):
        super().__init__()
        self.height_attn = TrajectoryAttention(d_model, n_heads, dropout)
        self.width_attn = TrajectoryAttention(d_model, n_heads, dropout)
        self._init_common(d_model, d_ffn, dropout, attn_drop, activation)
        self.return_attn_maps = False

    def forward(self, src: Tensor, pos: Tensor):
        if self.return_attn_maps:
            return self._forward_with_maps(src, pos)
        return self._run(src, pos), None, None

    def _forward_with_maps(self, src: Tensor, pos: Tensor):
        """Slow path for the attention visualiser: same kernels, one TrajectoryAttention at a time, maps materialised."""
        _require_inference(self, src, pos)
        B, T, H, W, C = pos.shape
        s0 = src.contiguous().float().view(B * T * H * W, C)
        p = pos.contiguous().float().view(B * T * H * W, C)
        pk = self.packed(src.device)
        s1 = ops.traj_attn_fwd(s0, s0, s0, p, s0, pk.attn_h, B, T, H, W, ops.AXIS_H)
        hmap = ops.traj_attn_maps(s0, p, pk.attn_h, B, T, H, W, ops.AXIS_H)
        s2 = ops.traj_attn_fwd(s1, s1, s1, p, s1, pk.attn_w, B, T, H, W, ops.AXIS_W)
        wmap = ops.traj_attn_maps(s1, p, pk.attn_w, B, T, H, W, ops.AXIS_W)
        out = ops.ln_ffn_fwd(s2, pk)
        return out.view(B * T, H * W, C).to(src.dtype), hmap, wmap


class TemporalEncoder(nn.Module):
    """WC/temporal_attention.py:79-100.  Note the reference quirk kept here: a `temporal_attn_type` that is neither
    "trajectory" nor "axial-trajectory" (e.g. the config default "axial_trajectory") creates no layers."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.0, attn_drop=0.0, activation="relu", n_heads=8,
                 temporal_attn_type="trajectory", num_temporal_layer=2):
        super().__init__()
        if temporal_attn_type == "trajectory":
            self.temporal_layers = nn.ModuleList([TemporalTrajectoryAttentionLayer(d_model, d_ffn, dropout, attn_drop, activation, n_heads)
                                                  for _ in range(num_temporal_layer)])
        elif temporal_attn_type == "axial-trajectory":
            self.temporal_layers = nn.ModuleList([TemporalAxialTrajectoryAttentionLayer(d_model, d_ffn, dropout, attn_drop, activation, n_heads)
                                                  for _ in range(num_temporal_layer)])

    def forward(self, src: Tensor, pos: Tensor):
        """src [B*T, H*W, C], pos [B, T, H, W, C] -> (src, height_traj_attn, width_traj_attn) of the last layer."""
        for layer in self.temporal_layers:
            src, height_traj_attn, width_traj_attn = layer(src, pos)
        return src, height_traj_attn, width_traj_attn

    def set_return_attn_maps(self, flag: bool = True) -> "TemporalEncoder":
        """Enable the slow path that materialises the attention maps of the LAST layer (what the reference returns)."""
        layers = list(self.temporal_layers)
        for i, layer in enumerate(layers):
            if hasattr(layer, "return_attn_maps"):
                layer.return_attn_maps = bool(flag) and i == len(layers) - 1
        return self
