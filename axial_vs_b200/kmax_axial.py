"""kMaX pixel-decoder axial attention.

Drop-ins for `AxialAttention` and `AxialAttention2D` of Vk/kmax_deeplab/modeling/pixel_decoder/kmax_pixel_decoder.py:104-190: same
constructor arguments, forward signatures and state-dict keys (`qkv_transform.conv.weight`, `_{query,key,value}_rpe._embeddings.weight`,
`_batch_norm_{qkv,similarity,retrieved_output}.*`; plain BatchNorm1d modules hold the SyncBatchNorm statistics), so `SingleBlock`
(:194-) can construct them unchanged.  Inference only: the batch norms apply their running statistics, folded into the GEMM weights /
the attention kernel's affines when the module is first run on a device.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib, ops
from .modules import _PackedCache, _invalidate_hook, _require_inference

MAX_SPAN = 255
_BN_EPS = 1e-3            # get_norm('syncbn'): eps = 1e-3, momentum = 0.01 (kmax_pixel_decoder.py:37)


class _Conv1x1(nn.Module):
    """`ConvBN(.., kernel_size=1, bias=False, norm=None, act=None, conv_type='1d')`: only the `conv.weight` key exists."""

    def __init__(self, c_in: int, c_out: int):
        super().__init__()
        self.conv = nn.Conv1d(c_in, c_out, kernel_size=1, bias=False)


class RelativePositionalEncoding(nn.Module):
    """kmax_pixel_decoder.py:89-102: an embedding table indexed by the relative distance m - l + MAX_SPAN - 1."""

    def __init__(self, query_length: int, key_length: int, depth: int):
        super().__init__()
        if query_length != key_length:
            raise NotImplementedError("axial_vs_b200: memory flange (key_length != query_length) is not used by any shipped config")
        self._embeddings = nn.Embedding(MAX_SPAN * 2 - 1, depth)
        nn.init.trunc_normal_(self._embeddings.weight, std=1.0, a=-2.0, b=2.0)
        self.query_length, self.key_length, self.depth = query_length, key_length, depth

    def forward(self):
        idx = torch.arange(self.key_length)[None, :] - torch.arange(self.query_length)[:, None] + MAX_SPAN - 1
        return self._embeddings.weight[idx.to(self._embeddings.weight.device)]


def _fold(bn: nn.BatchNorm1d):
    s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    return s, bn.bias.detach().double() - bn.running_mean.detach().double() * s


class AxialAttention(nn.Module):
    """forward(x [N, in_planes, L]) -> [N, total_value_depth, L]  (kmax_pixel_decoder.py:105-157)."""

    def __init__(self, in_planes, query_shape=56, total_key_depth=512, total_value_depth=1024, num_heads=8):
        assert (total_key_depth % num_heads == 0) and (total_value_depth % num_heads == 0)
        super().__init__()
        self._in_planes, self._query_shape = in_planes, query_shape
        self._total_key_depth, self._total_value_depth, self._num_heads = total_key_depth, total_value_depth, num_heads
        self._key_depth_per_head = total_key_depth // num_heads
        n_qkv = total_key_depth * 2 + total_value_depth
        self.qkv_transform = _Conv1x1(in_planes, n_qkv)
        nn.init.trunc_normal_(self.qkv_transform.conv.weight, std=in_planes ** -0.5)
        self._query_rpe = RelativePositionalEncoding(query_shape, query_shape, self._key_depth_per_head)
        self._key_rpe = RelativePositionalEncoding(query_shape, query_shape, self._key_depth_per_head)
        self._value_rpe = RelativePositionalEncoding(query_shape, query_shape, total_value_depth // num_heads)
        self._batch_norm_qkv = nn.BatchNorm1d(n_qkv, eps=_BN_EPS, momentum=0.01)
        self._batch_norm_similarity = nn.BatchNorm1d(num_heads * 3, eps=_BN_EPS, momentum=0.01)
        self._batch_norm_retrieved_output = nn.BatchNorm1d(total_value_depth * 2, eps=_BN_EPS, momentum=0.01)
        self._cache = _PackedCache()
        self.register_load_state_dict_post_hook(_invalidate_hook)

    def _packed(self, device):
        def build():
            s, t = _fold(self._batch_norm_qkv)
            w = (self.qkv_transform.conv.weight.detach().double()[:, :, 0] * s[:, None]).float().to(device)
            w = torch.cat((w, w, w - w.bfloat16().float()), dim=1).contiguous()      # split precision: [W_hi | W_hi | W_lo] (include/axvs.h)
            f32 = lambda v: v.detach().float().contiguous().to(device)
            ss, st = _fold(self._batch_norm_similarity)
            os_, ot = _fold(self._batch_norm_retrieved_output)
            return {"w_qkv": ops.pack_weight(w), "b_qkv": f32(t), "emb_q": f32(self._query_rpe._embeddings.weight),
                    "emb_k": f32(self._key_rpe._embeddings.weight), "emb_v": f32(self._value_rpe._embeddings.weight),
                    "sim_s": f32(ss), "sim_t": f32(st), "out_s": f32(os_), "out_t": f32(ot)}
        return self._cache.get(self, device, build)

    def _struct(self, pk) -> _lib.KmaxAxialWeights:
        return _lib.KmaxAxialWeights(*(pk[k].data_ptr() for k in ("w_qkv", "b_qkv", "emb_q", "emb_k", "emb_v", "sim_s", "sim_t", "out_s", "out_t")),
                                     self._num_heads, self._key_depth_per_head, self._total_value_depth // self._num_heads)

    def run(self, x: torch.Tensor, x_layout: int, images: int, H: int, W: int, axis: int, out_layout: int) -> torch.Tensor:
        """One pass through the C ABI (`axvs_kmax_axial_fwd`); x fp32 NCHW (x_layout 0) or token rows (1)."""
        if x.device.type != "cuda":
            raise RuntimeError("axial_vs_b200: CUDA tensors required (there is no CPU fallback)")
        pk = self._packed(x.device)
        Vd, heads = self._total_value_depth, self._num_heads
        out = torch.empty((images, Vd, H, W) if out_layout == 0 else (images * H * W, Vd), dtype=torch.float32, device=x.device)
        lib = _lib.load()
        nbytes = lib.axvs_kmax_axial_workspace_bytes(images, self._in_planes, H, W, heads, self._key_depth_per_head, Vd // heads)
        with torch.cuda.device(x.device):
            ws = ops.workspace(nbytes, x.device)
            st = self._struct(pk)
            rc = lib.axvs_kmax_axial_fwd(x.data_ptr(), x_layout, images, self._in_planes, H, W, axis, ctypes.byref(st), out.data_ptr(), out_layout,
                                         ws.data_ptr(), ws.numel(), ops._stream(x.device))
        _lib.check(rc, "axvs_kmax_axial_fwd")
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_inference(self, x)
        N, C, L = x.shape
        if C != self._in_planes:
            raise RuntimeError(f"expected {self._in_planes} input channels, got {C}")
        y = self.run(x.contiguous().float(), 0, N, L, 1, 1, 0)                     # [N, Vd, L, 1]
        return y.view(N, self._total_value_depth, L).to(x.dtype)


class AxialAttention2D(nn.Module):
    """forward(x [N, in_planes, H, W]) -> [N, total_value_depth, H, W]  (kmax_pixel_decoder.py:161-190).  The reference's two
    permute + contiguous copies are folded into the kernels' addressing: the height pass writes token rows, the width pass NCHW."""

    def __init__(self, in_planes, query_shape=[56, 56], filters=512, key_expansion=1, value_expansion=2, num_heads=8):
        super().__init__()
        total_key_depth = int(round(filters * key_expansion))
        total_value_depth = int(round(filters * value_expansion))
        self._total_key_depth, self._total_value_depth = total_key_depth, total_value_depth
        self._height_axis = AxialAttention(in_planes, query_shape[0], total_key_depth, total_value_depth, num_heads)
        self._width_axis = AxialAttention(total_value_depth, query_shape[1], total_key_depth, total_value_depth, num_heads)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_inference(self, x)
        N, C, H, W = x.shape
        rows = self._height_axis.run(x.contiguous().float(), 0, N, H, W, 1, 1)    # token rows [(n h w), Vd]
        y = self._width_axis.run(rows, 1, N, H, W, 2, 0)                          # NCHW
        return y.to(x.dtype)
