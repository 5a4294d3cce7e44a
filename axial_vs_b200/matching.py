"""Clip-to-clip query matching on the GPU.

Drop-ins for `MaXTronWCDeepLab.match_from_embds` (Vk/maxtron_deeplab/maxtron_wc_model.py:391-400; the cross-clip model carries a copy) and
for the chains around it (maxtron_wc_model.py:342-346, maxtron_cc_model.py:280-298).  The reference moves a 128 x 128 cost matrix to the
host for `scipy.optimize.linear_sum_assignment` once per adjacent clip pair; here the cosine cost, the exact assignment (scipy's own
algorithm and tie rules, include/axvs.h) and the chaining run on the device, one launch per chain, no host synchronisation.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import _lib, ops


def _cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.device.type != "cuda":
        raise RuntimeError(f"axial_vs_b200: {name} must be a CUDA tensor (there is no CPU fallback)")
    return t.contiguous().float()


@torch.no_grad()
def linear_sum_assignment(cost: torch.Tensor) -> torch.Tensor:
    """cost [n, n] or [batch, n, n] (row = target, column = current), n <= 256 -> int64 column indices like
    `scipy.optimize.linear_sum_assignment(cost)[1]` (per matrix), on the device."""
    c = _cuda(cost, "cost")
    squeeze = c.dim() == 2
    if squeeze:
        c = c.unsqueeze(0)
    if c.dim() != 3 or c.shape[1] != c.shape[2]:
        raise RuntimeError(f"linear_sum_assignment: square matrices expected, got {tuple(cost.shape)}")
    b, n, _ = c.shape
    out = torch.empty(b, n, dtype=torch.int32, device=c.device)
    lib = _lib.load()
    with torch.cuda.device(c.device):
        _lib.check(lib.axvs_lsap(c.data_ptr(), b, n, out.data_ptr(), ops._stream(c.device)), "axvs_lsap")
    out = out.long()
    return out[0] if squeeze else out


@torch.no_grad()
def match_chain(embeddings: torch.Tensor) -> torch.Tensor:
    """embeddings [clips, n, e] or [videos, clips, n, e] -> int64 indices of the same leading shape + [n]: row 0 is the identity and
    `embeddings[i][indices[i]]` is clip i aligned to clip 0 through the chain of pairwise matches (maxtron_wc_model.py:342-346)."""
    x = _cuda(embeddings, "embeddings")
    squeeze = x.dim() == 3
    if squeeze:
        x = x.unsqueeze(0)
    if x.dim() != 4:
        raise RuntimeError(f"match_chain: [clips, n, e] or [videos, clips, n, e] expected, got {tuple(embeddings.shape)}")
    v, clips, n, e = x.shape
    out = torch.empty(v, clips, n, dtype=torch.int32, device=x.device)
    lib = _lib.load()
    nbytes = lib.axvs_match_chain_workspace_bytes(v, n, e)
    with torch.cuda.device(x.device):
        ws = ops.workspace(nbytes, x.device)
        _lib.check(lib.axvs_match_chain(x.data_ptr(), v, clips, n, e, out.data_ptr(), ws.data_ptr(), ws.numel(), ops._stream(x.device)),
                   "axvs_match_chain")
    out = out.long()
    return out[0] if squeeze else out


@torch.no_grad()
def match_from_embds(tgt_embds: torch.Tensor, cur_embds: torch.Tensor) -> torch.Tensor:
    """Reference signature (maxtron_wc_model.py:391): the permutation that aligns `cur_embds` [n, e] to `tgt_embds` [n, e]."""
    return match_chain(torch.stack((tgt_embds, cur_embds), 0))[1]


@torch.no_grad()
def align_clips(mask_embeddings: Sequence[torch.Tensor], *others: Sequence[torch.Tensor]) -> List[List[torch.Tensor]]:
    """The reference's video-wise matching loop (maxtron_wc_model.py:337-346): `mask_embeddings[i]` [n, e] per clip; every sequence in
    `others` holds per-clip tensors whose FIRST dimension is the query dimension (pred_masks [n, T, H, W], pred_logits [n, K], cluster
    centres ...).  Returns [aligned mask embeddings, aligned others...]; everything stays on the device."""
    idx = match_chain(torch.stack(list(mask_embeddings), 0))
    out = [[m[idx[i]] for i, m in enumerate(mask_embeddings)]]
    for seq in others:
        out.append([t[idx[i]] for i, t in enumerate(seq)])
    return out
