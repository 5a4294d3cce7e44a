"""The assignment algorithm the CUDA matching kernel implements (oracle/matching_oracle.lsap, a restatement of scipy's
rectangular_lsap.cpp) against scipy.optimize.linear_sum_assignment ITSELF -- the reference's own dependency for this step
(Vk/maxtron_deeplab/maxtron_wc_model.py:398) -- including heavily tied matrices, where only the same tie rules give the same permutation."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from oracle import matching_oracle as M


def _cases():
    rng = np.random.default_rng(0)
    out = []
    for n in (1, 2, 3, 5, 8, 17, 33, 64):
        out.append(("uniform", rng.random((n, n)).astype(np.float32)))
        out.append(("ties_small_int", rng.integers(0, 3, (n, n)).astype(np.float32)))
        out.append(("ties_int", rng.integers(0, 10, (n, n)).astype(np.float32)))
        out.append(("constant", np.full((n, n), 0.25, dtype=np.float32)))
        out.append(("rank1", np.outer(np.arange(n), np.arange(n)).astype(np.float32)))
        out.append(("negative", (rng.random((n, n)) - 0.5).astype(np.float32)))
        b = rng.integers(0, 2, (n, n)).astype(np.float32)
        out.append(("binary", b))
    e = rng.standard_normal((24, 16)).astype(np.float32)                # cosine cost with duplicated rows (exact ties)
    e[5] = e[3]; e[11] = e[3]
    en = e / np.linalg.norm(e, axis=1, keepdims=True)
    out.append(("cosine_dup", (1 - en @ en[rng.permutation(24)].T).astype(np.float32)))
    return out


@pytest.mark.parametrize("name,cost", _cases(), ids=lambda x: x if isinstance(x, str) else f"n{x.shape[0]}")
def test_lsap_restatement_matches_scipy(name, cost):
    want = linear_sum_assignment(cost)[1]
    got = M.lsap(cost)
    assert np.array_equal(got, want), (name, cost.shape)


def test_lsap_infeasible():
    c = np.full((3, 3), np.inf, dtype=np.float32)
    assert (M.lsap(c) == -1).all()
    with pytest.raises(ValueError):
        linear_sum_assignment(c)


def test_chain_is_a_composition_of_pairwise_matches():
    g = torch.Generator().manual_seed(3)
    base = torch.randn(12, 32, generator=g)
    clips = [base]
    perms = []
    for i in range(4):                                                     # every clip = a permuted, slightly perturbed copy of the previous one
        p = torch.randperm(12, generator=g)
        perms.append(p)
        clips.append(clips[-1][p] + 0.01 * torch.randn(12, 32, generator=g))
    idx = M.match_chain(torch.stack(clips, 0))
    aligned = [clips[i][torch.as_tensor(idx[i])] for i in range(5)]
    for a in aligned[1:]:
        assert torch.nn.functional.cosine_similarity(a, base, dim=1).min() > 0.99   # every clip lands on clip 0's query order
