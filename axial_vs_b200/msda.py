"""MSDeformAttn spatial encoder layer of the within-clip tracking module.

Drop-in for `MSDeformAttnTransformerEncoderLayer` (`WC/msdeformattn.py:177-215`) with its `MSDeformAttn` sub-module
(`WC/ops/modules/ms_deform_attn.py:34-125`): same constructor arguments, forward signature and state-dict keys
(`self_attn.{sampling_offsets,attention_weights,value_proj,output_proj}.*`, `norm1`, `linear1`, `linear2`, `norm2`).
Inference only, no padding (the reference's module passes all-False masks, `WC/msdeformattn.py:92`).  The four Linear layers
and the FFN run on the tcgen05 kernels, the multi-scale bilinear gather in `msda_sample_kernel`.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import ops
from .modules import _PackedCache, _invalidate_hook, _require_inference


class MSDeformAttn(nn.Module):
    """Parameter holder with the reference's names and initialisation (the arithmetic is fused into the layer call)."""

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model != ops.C or n_heads != ops.HEADS:
            raise NotImplementedError(f"axial_vs_b200 kernels are specialised for d_model=256, n_heads=8 (got {d_model}, {n_heads})")
        if n_levels > 4 or n_levels * n_points > 16:
            raise NotImplementedError("axial_vs_b200: at most 4 levels and 16 level*point samples per head")
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.im2col_step = 128
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        """Same initial values as the reference (WC/ops/modules/ms_deform_attn.py:67-83): zero offset / attention weights, offset
        biases pointing in 8 compass directions (one per head, unit max-norm) scaled by the point index 1..P, Xavier projections."""
        ang = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        dirs = torch.stack([ang.cos(), ang.sin()], dim=-1)
        dirs = dirs / dirs.abs().amax(dim=-1, keepdim=True)                                  # [heads, 2]
        steps = torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, self.n_points, 1)
        bias = (dirs.view(self.n_heads, 1, 1, 2) * steps).expand(-1, self.n_levels, -1, -1)  # [heads, levels, points, 2]
        with torch.no_grad():
            self.sampling_offsets.weight.zero_()
            self.sampling_offsets.bias.copy_(bias.reshape(-1))
            self.attention_weights.weight.zero_()
            self.attention_weights.bias.zero_()
            for lin in (self.value_proj, self.output_proj):
                nn.init.xavier_uniform_(lin.weight)
                lin.bias.zero_()


class MSDeformAttnTransformerEncoderLayer(nn.Module):
    """forward(src [images, len, 256], pos, reference_points [images, len, n_levels, 2], spatial_shapes, level_start_index,
    padding_mask=None) -> src."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if activation != "relu":
            raise NotImplementedError("axial_vs_b200: only activation='relu' is implemented (every shipped config uses it)")
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self._cache = _PackedCache()
        self.register_load_state_dict_post_hook(_invalidate_hook)

    def packed(self, device) -> ops.PackedMsda:
        return self._cache.get(self, device, lambda: ops.pack_msda_layer(dict(self.state_dict()), self.self_attn.n_levels, self.self_attn.n_points))

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index=None, padding_mask=None):
        _require_inference(self, src)
        if padding_mask is not None and bool(padding_mask.any()):
            raise NotImplementedError("axial_vs_b200: padded feature maps are not supported (the reference module never pads)")
        if reference_points.shape[-1] != 2:
            raise NotImplementedError("axial_vs_b200: 2-D reference points only (encoder use)")
        shapes = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes) else spatial_shapes)]
        out = ops.msda_layer_fwd(src.contiguous().float(), None if pos is None else pos.contiguous().float(),
                                 reference_points.contiguous().float(), shapes, self.packed(src.device))
        return out.to(src.dtype)


def reference_points(spatial_shapes, images: int, device) -> torch.Tensor:
    """`MSDeformAttnTransformerEncoder.get_reference_points` for unpadded maps (valid_ratios == 1), WC/msdeformattn.py:231-245."""
    pts = []
    for (H, W) in spatial_shapes:
        ys, xs = torch.meshgrid(torch.linspace(0.5, H - 0.5, H, device=device), torch.linspace(0.5, W - 0.5, W, device=device), indexing="ij")
        pts.append(torch.stack((xs.reshape(-1) / W, ys.reshape(-1) / H), -1))
    ref = torch.cat(pts, 0)
    return ref[None, :, None, :].expand(images, -1, len(spatial_shapes), -1).contiguous()
