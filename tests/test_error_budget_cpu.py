"""Error budget of the final per-pixel labels (BASELINE.json north_star: >= 99.9 % per-pixel agreement on the final argmax labels).

The cross-clip module ends in `argmax_q mask_logits[q, pixel]` over 128 queries.  With random-init heads those logits are nearly tied
(median top-2 margin ~ 3 % of the logit range), so the label agreement measures operand rounding directly.  This test emulates the
CUDA path's bf16 operand rounding STAGE BY STAGE inside the fp32 oracle (oracle/traj_oracle.EMULATE_BF16) and compares the labels with
the float64 reference, on a cfg3-shaped shard (Q = 128 queries, 4 cross-clip layers).  What it pins:

  * every stage after the trajectory attention (temporal ASPP, ConvBN projections, query x pixel contraction) costs MORE label
    agreement in bf16 than the whole bf16 trajectory attention does -> those stages run split-precision (fp32-grade) in the CUDA path
    (csrc/cc_tail.cuh, axvs_linear_f32, axvs_mask_einsum_f32); round 1 ran them in bf16 and measured 99.0 %;
  * with only the trajectory attention in bf16 (what the CUDA path does now) the agreement is ~99.85 %: the remaining 0.1-0.2 % of the
    pixels have a reference top-2 margin below the bf16-compute error of the attention itself, which north_star prescribes
    ("bf16 compute with fp32 accumulation").  The GPU test `test_label_agreement_end_to_end` asserts the measured figure.
"""
import torch

from axial_vs_b200 import synth
from oracle import traj_oracle as O


def _agreement(stages, cq, pf, p, L, V, ref):
    O.EMULATE_BF16 = set(stages)
    try:
        out = O.cross_clip_module(cq, pf, p, L, V)["pred_masks"].double()
    finally:
        O.EMULATE_BF16 = set()
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    return (out.argmax(1) == ref.argmax(1)).float().mean().item(), err


def test_label_error_budget_by_stage(capsys):
    Q, T, V, H, W, L, K, seed = 128, 6, 2, 20, 20, 4, 124, 909
    p = synth.cross_clip_params(seed, L, K)
    cq = synth.randn(seed + 1, 1, Q, T, 256)
    pf = synth.randn(seed + 2, 1, 128, T * V, H, W)
    ref = O.cross_clip_module(cq.double(), pf.double(), {k: v.double() for k, v in p.items()}, L, V)["pred_masks"]
    top2 = ref.topk(2, dim=1).values
    rel_margin = ((top2[:, 0] - top2[:, 1]).median() / ref.abs().max()).item()
    rows = {}
    for stages in [(), ("ta",), ("aspp",), ("proj",), ("einsum",), ("aspp", "proj", "einsum"), ("ta", "aspp", "proj", "einsum")]:
        rows[stages] = _agreement(stages, cq, pf, p, L, V, ref)
    with capsys.disabled():
        print(f"\n[label error budget] median top-2 margin = {rel_margin:.3%} of the logit range")
        for stages, (agree, err) in rows.items():
            print(f"  bf16 operands in {'+'.join(stages) or 'no stage (fp32)':28s} label agreement {agree:.4%}   max-normalised logit error {err:.1e}")
    assert rows[()][0] >= 0.9999                                   # fp32 arithmetic reproduces the float64 labels
    ta = rows[("ta",)][0]
    for s in ("aspp", "proj", "einsum"):
        assert rows[(s,)][0] < ta, f"stage {s} was expected to cost more label agreement than the trajectory attention"
    assert ta >= 0.997                                             # what the CUDA path (attention in bf16, the rest fp32-grade) can reach
    assert rows[("ta", "aspp", "proj", "einsum")][0] < 0.995       # round 1's all-bf16 arithmetic: ~99.1 %
    assert ta - rows[("ta", "aspp", "proj", "einsum")][0] > 0.004  # the split-precision tail buys > 0.4 % of the pixels


def test_emulation_is_off_by_default():
    assert O.EMULATE_BF16 == set()
