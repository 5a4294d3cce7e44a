"""Within-clip input / output projections.

Drop-ins for the `nn.Sequential(nn.Conv2d(.., kernel_size=1), nn.GroupNorm(32, ..))` pairs of
`MSDeformAttnPixelDecoder` (`WC/msdeformattn.py:355-375`): same parameter names (`0.weight`, `0.bias`, `1.weight`,
`1.bias`), so `input_proj.{i}.*` / `output_proj.{i}.*` checkpoint entries load unchanged.  The input side returns
TOKEN-MAJOR rows `[images, H*W, 256]` -- what `MSDeformAttnTransformerEncoderOnly` builds with `flatten(2).transpose(1, 2)`
(`WC/msdeformattn.py:100-106`) and what the temporal layers consume; the output side takes token rows and returns NCHW
(the reference's `transpose(1, 2).view(bs, -1, H, W)` at `:432-434` is folded into the GEMM's store).  Inference only.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .modules import _PackedCache, _invalidate_hook, _require_inference


class _Proj(nn.Sequential):
    def __init__(self, c_in: int, c_out: int):
        super().__init__(nn.Conv2d(c_in, c_out, kernel_size=1), nn.GroupNorm(32, c_out))
        nn.init.xavier_uniform_(self[0].weight, gain=1)      # WC/msdeformattn.py:377-382
        nn.init.constant_(self[0].bias, 0)
        self._cache = _PackedCache()
        self.register_load_state_dict_post_hook(_invalidate_hook)

    def _packed(self, device):
        def build():
            w = self[0].weight.detach().float().flatten(1)
            n_pad = (w.shape[0] + 255) // 256 * 256                 # output side: the GEMM stores 256-column chunks (padding never stored)
            if n_pad != w.shape[0]:
                w = torch.cat((w, w.new_zeros(n_pad - w.shape[0], w.shape[1])), 0)
            return ops.pack_weight(w.contiguous())
        return self._cache.get(self, device, build)

    def _params(self):
        b = self[0].bias.detach().float()
        n_pad = (b.numel() + 255) // 256 * 256
        if n_pad != b.numel():
            b = torch.cat((b, b.new_zeros(n_pad - b.numel())))
        return (b.contiguous(), self[1].weight.detach().float().contiguous(), self[1].bias.detach().float().contiguous())


class InputProjection(_Proj):
    """forward(x [images, c_in, H, W]) -> tokens [images, H*W, 256] (c_in a multiple of 64)."""

    def __init__(self, in_channels: int, conv_dims: int = 256):
        if conv_dims != 256:
            raise NotImplementedError("axial_vs_b200: conv_dims must be 256 (every shipped config)")
        super().__init__(in_channels, conv_dims)

    def forward(self, x: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """`out` (optional, fp32): the level's slice of the multi-level token tensor; the projection writes into it (no torch.cat afterwards)."""
        _require_inference(self, x)
        b, gw, gb = self._params()
        y = ops.input_proj_fwd(x.contiguous().float(), self._packed(x.device), b, gw, gb, self[1].eps, out=out)
        return y if out is not None else y.to(x.dtype)


class OutputProjection(_Proj):
    """forward(tokens [images, H*W, 256], H, W) -> y [images, c_out, H, W] (c_out a multiple of 32, e.g. 384 / 768 / 1536 for ConvNeXt-L)."""

    def __init__(self, out_channels: int, conv_dims: int = 256):
        if conv_dims != 256:
            raise NotImplementedError("axial_vs_b200: conv_dims must be 256 (every shipped config)")
        super().__init__(conv_dims, out_channels)

    def forward(self, tokens: torch.Tensor, H: int, W: int) -> torch.Tensor:
        _require_inference(self, tokens)
        b, gw, gb = self._params()
        t = tokens.float()
        if not (t.stride(2) == 1 and t.stride(1) == ops.C and t.data_ptr() % 16 == 0):       # dense rows or a level slice are read in place
            t = t.contiguous()
        return ops.output_proj_fwd(t, self._packed(tokens.device), b, gw, gb, H, W, self[1].eps).to(tokens.dtype)
