"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's clip-to-clip query matching.

`match_from_embds` / `match_chain` follow MaXTronWCDeepLab.match_from_embds (Vk/maxtron_deeplab/maxtron_wc_model.py:391-400) and the loop
at :337-346 (copy: maxtron_cc_model.py:280-298), calling the reference's own dependency `scipy.optimize.linear_sum_assignment` (scipy is
present in this image; the reference pins no version).

`lsap` restates the algorithm behind that call -- scipy/optimize/rectangular_lsap/rectangular_lsap.cpp (Crouse's shortest augmenting path
implementation of Jonker-Volgenant), which is not vendored under /root/reference -- in plain Python, float64, with scipy's tie rules: this
is what the CUDA kernel implements (csrc/matching.cuh).  Pinned: tests/test_matching_cpu.py checks `lsap` against scipy itself on random,
integer-valued (heavily tied), constant and adversarial matrices, so the restatement (and through it the kernel) is anchored on the real thing.
Only tests/, `__graft_entry__.smoke()` and bench.py's CPU legs may import this file.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def lsap(cost: np.ndarray) -> np.ndarray:
    """col4row of the square float matrix `cost` (row = target); -1 everywhere if infeasible."""
    c = np.asarray(cost, dtype=np.float64)
    n = c.shape[0]
    assert c.shape == (n, n)
    u, v = np.zeros(n), np.zeros(n)
    col4row, row4col, path = [-1] * n, [-1] * n, [-1] * n
    for cur_row in range(n):
        remaining = [n - it - 1 for it in range(n)]              # filled in reverse: a constant matrix yields the identity
        SR, SC = [False] * n, [False] * n
        spc = [math.inf] * n
        num_remaining, sink, min_val, i = n, -1, 0.0, cur_row
        while sink == -1:
            index, lowest = -1, math.inf
            SR[i] = True
            for it in range(num_remaining):
                j = remaining[it]
                r = min_val + c[i, j] - u[i] - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    index = it
            min_val = lowest
            if min_val == math.inf:
                return np.full(n, -1, dtype=np.int64)
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            num_remaining -= 1
            remaining[index] = remaining[num_remaining]
        u[cur_row] += min_val
        for k in range(n):
            if SR[k] and k != cur_row:
                u[k] += min_val - spc[col4row[k]]
        for j in range(n):
            if SC[j]:
                v[j] -= min_val - spc[j]
        j = sink
        while True:
            k = path[j]
            row4col[j] = k
            col4row[k], j = j, col4row[k]
            if k == cur_row:
                break
    return np.asarray(col4row, dtype=np.int64)


def match_from_embds(tgt_embds: torch.Tensor, cur_embds: torch.Tensor) -> np.ndarray:
    """maxtron_wc_model.py:391-400."""
    from scipy.optimize import linear_sum_assignment
    cur = cur_embds / cur_embds.norm(dim=1)[:, None]
    tgt = tgt_embds / tgt_embds.norm(dim=1)[:, None]
    C = (1 - torch.mm(cur, tgt.transpose(0, 1))).cpu()
    return linear_sum_assignment(C.transpose(0, 1))[1]           # target x current -> permutation of current


def match_chain(embeddings: torch.Tensor) -> np.ndarray:
    """maxtron_wc_model.py:337-346 for one video: embeddings [clips, n, e] -> indices [clips, n] (row 0 = identity)."""
    clips, n, _ = embeddings.shape
    out = [np.arange(n)]
    prev = embeddings[0]
    for i in range(1, clips):
        idx = match_from_embds(prev, embeddings[i])
        out.append(np.asarray(idx))
        prev = embeddings[i][torch.as_tensor(idx)]
    return np.stack(out, 0)
