"""CPU: the oracle restatement against the committed golden fixtures (reference outputs, fp32)."""
import numpy as np
import pytest
import torch

from axial_vs_b200 import synth
from oracle import traj_oracle as O

ATOL = 2e-5  # fp32 restatement vs fp32 reference: different summation order only


def _close(a, b, atol=ATOL):
    a = torch.as_tensor(np.asarray(a)).float()
    b = torch.as_tensor(np.asarray(b)).float()
    assert a.shape == b.shape
    err = (a - b).abs().max().item()
    assert err <= atol * max(1.0, b.abs().max().item()), f"max abs err {err}"


@pytest.mark.parametrize("tag", ["a", "b"])
def test_trajectory_attention(golden, tag):
    gz = golden(f"ta_vk_{tag}")
    Bp, F, n, seed = int(gz["Bp"]), int(gz["F"]), int(gz["n"]), int(gz["seed"])
    p = {}
    synth.traj_attn_params(torch.Generator().manual_seed(seed), "", 256, p)
    assert synth.checksum(p) == pytest.approx(float(gz["wsum"]), rel=1e-12)
    q = synth.randn(seed + 100, Bp, F * n, 256)
    v = synth.randn(seed + 200, Bp, F * n, 256)
    y, maps = O.trajectory_attention(q, q, v, p, F, return_maps=True)
    _close(y, gz["y"])
    _close(maps, gz["maps"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_axial_layer(golden, tag):
    gz = golden(f"axial_layer_{tag}")
    B, T, H, W, seed = (int(gz[k]) for k in "B T H W seed".split())
    p = synth.axial_layer_params(seed)
    assert synth.checksum(p) == pytest.approx(float(gz["wsum"]), rel=1e-12)
    src = synth.randn(seed + 100, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 200)[0])
    out, hm, wm = O.axial_layer(src, pos, p, return_maps=True)
    _close(out, gz["out"])
    _close(hm, gz["hmap"])
    _close(wm, gz["wmap"])


def test_encoder_axial(golden):
    gz = golden("encoder_axial")
    B, T, H, W, seed = (int(gz[k]) for k in "B T H W seed".split())
    p = synth.encoder_params(seed, 2)
    src = synth.randn(seed + 100, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 200)[1])
    out, hm, wm = O.temporal_encoder(src, pos, O.split_encoder_params(p), return_maps=True)
    _close(out, gz["out"])
    _close(hm, gz["hmap"])
    _close(wm, gz["wmap"])


def test_encoder_trajectory(golden):
    gz = golden("encoder_trajectory")
    B, T, H, W, seed = (int(gz[k]) for k in "B T H W seed".split())
    p = synth.encoder_params(seed, 1, axial=False)
    src = synth.randn(seed + 100, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 200)[0])
    out, _, _ = O.temporal_encoder(src, pos, O.split_encoder_params(p), attn_type="trajectory")
    _close(out, gz["out"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_pos3d(golden, tag):
    gz = golden(f"pos3d_{tag}")
    B, T, H, W = (int(gz[k]) for k in "B T H W".split())
    _close(O.pos3d_table(B, T, H, W), gz["table"], atol=1e-6)


def test_cc_trajectory_attention(golden):
    gz = golden("cc_ta")
    b, Q, T, seed = (int(gz[k]) for k in "b Q T seed".split())
    p = {}
    synth.traj_attn_params(torch.Generator().manual_seed(seed), "", 256, p, fused_qkv=True)
    x = synth.randn(seed + 100, b, T * Q, 256)
    _close(O.cc_trajectory_attention(x, p, Q, T), gz["y"])


def test_cc_module(golden):
    gz = golden("cc_module")
    Q, T, V, H, W, L, K, seed = (int(gz[k]) for k in "Q T V H W L K seed".split())
    p = synth.cross_clip_params(seed, L, K)
    assert synth.checksum(p) == pytest.approx(float(gz["wsum"]), rel=1e-12)
    cq = synth.randn(seed + 100, 1, Q, T, 256)
    pf = synth.randn(seed + 200, 1, 128, T * V, H, W)
    o = O.cross_clip_module(cq, pf, p, L, V)
    _close(o["pred_logits"], gz["pred_logits"])
    _close(o["pred_masks"], gz["pred_masks"], atol=5e-5)


def test_fp64_oracle_agrees_with_fp32():
    """The oracle is dtype-generic: fp64 run bounds the fp32 reference's own rounding noise."""
    p = synth.axial_layer_params(5)
    src = synth.randn(6, 2, 12, 256)
    pos = O.level_pos3d(1, 2, 3, 4, synth.level_embed(7)[0])
    o32, _, _ = O.axial_layer(src, pos, p)
    o64, _, _ = O.axial_layer(src.double(), pos.double(), p)
    assert (o32.double() - o64).abs().max().item() < 2e-5


def test_flop_formulas_match_baseline_md():
    assert abs(O.flops_axial_layer(1, 2, 41, 41) / 1e9 - 12.04) < 0.01
    assert abs(O.flops_axial_layer(1, 2, 21, 21) / 1e9 - 3.09) < 0.01
    assert abs(O.flops_axial_layer(1, 10, 161, 161) / 1e9 - 2830.56) < 0.1


@pytest.mark.parametrize("tag", ["a", "b"])
def test_decoder_attention(golden, tag):
    """Row A11: AttentionOperation and the k-means update against tensors captured inside the reference kMaXTransformerLayer."""
    gz = golden(f"decoder_attn_{tag}")
    bn = {k[3:]: torch.as_tensor(gz[k]) for k in gz.files if k.startswith("bn.")}
    t = lambda k: torch.as_tensor(gz[k])
    _close(O.query_self_attention(t("q"), t("k"), t("v"), bn), gz["attn_out"])
    upd, idx = O.kmeans_update(t("mask_logits"), t("pixel_value"))
    _close(upd, gz["kmeans_update"])
    assert torch.equal(idx, t("mask_logits").argmax(1))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_projections(golden, tag):
    """Row f1: Conv2d 1x1 + GroupNorm(32) input / output projections against the stock torch modules the reference instantiates."""
    gz = golden(f"proj_{tag}")
    n, c, H, W, seed = (int(gz[k]) for k in "n c H W seed".split())
    pin, pout = synth.proj_params(seed, c)
    assert synth.checksum({**{"i." + k: v for k, v in pin.items()}, **{"o." + k: v for k, v in pout.items()}}) == pytest.approx(float(gz["wsum"]), rel=1e-12)
    x = synth.randn(seed + 100, n, c, H, W)
    tok = O.input_proj(x, pin)
    _close(tok, gz["tokens"])
    _close(O.output_proj(tok, pout, H, W), gz["y"], atol=5e-5)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_msda_layer(golden, tag):
    """Row f2: MSDeformAttn spatial encoder layer against the reference module run on CPU (its pure-PyTorch sampling branch)."""
    gz = golden(f"msda_layer_{tag}")
    n, seed = int(gz["n"]), int(gz["seed"])
    shapes = [tuple(int(v) for v in r) for r in gz["shapes"]]
    p = synth.msda_layer_params(seed)
    assert synth.checksum(p) == pytest.approx(float(gz["wsum"]), rel=1e-12)
    Len = sum(h * w for h, w in shapes)
    src, pos = synth.randn(seed + 100, n, Len, 256), synth.randn(seed + 200, n, Len, 256)
    ref = O.msda_reference_points(shapes, n)
    _close(ref, gz["ref_points"])
    _close(O.msda_encoder_layer(src, pos, ref, shapes, p), gz["out"])


from _cases import wc_encoder_case as _wc_encoder_case  # noqa: E402


def test_within_clip_encoder(golden):
    """Rows A6 + f2 together: the reference MSDeformAttnTransformerEncoder (2 stages, spatial layer on 3 levels + temporal layer
    on 2) against the oracle's composition of its spatial and temporal restatements."""
    gz = golden("wc_encoder")
    B, T, shapes, spatial, temporal_states, state, src, pos, pos3d = _wc_encoder_case(gz)
    assert synth.checksum(state) == pytest.approx(float(gz["wsum"]), rel=1e-12)
    temporal = [O.split_encoder_params(ts) for ts in temporal_states]
    out = O.within_clip_encoder(src, shapes, pos, pos3d, spatial, temporal, 2)
    _close(out, gz["out"], atol=5e-5)


def test_within_clip_module(golden):
    """The whole within-clip tracking module (MSDeformAttnPixelDecoder.forward_features of the reference, CPU): projections, 2-D and
    3-D positional terms, two stages of spatial + temporal layers, output projections."""
    gz = golden("wc_module")
    seed, chans = int(gz["seed"]), [int(c) for c in gz["chans"]]
    sizes = [tuple(int(v) for v in r) for r in gz["sizes"]]
    p = synth.within_clip_module_params(seed, chans)
    assert synth.checksum(p) == pytest.approx(float(gz["wsum"]), rel=1e-12)
    feats = [synth.randn(seed + 100 + i, 2, chans[i], *sizes[i]) for i in range(3)]
    outs = O.within_clip_module(feats, p, 1, 2)
    for o, name in zip(outs, ("res5", "res4", "res3")):
        _close(o, gz[name], atol=2e-4)


# --------------------------------------------------------------------------------------------- post-path tail (row f4)
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_panoptic_mask_inference(golden, tag):
    """The numpy restatement against outputs of the unmodified `MaXTronWCDeepLab.panoptic_mask_inference` (integer ids: exact)."""
    from oracle import panoptic_oracle as PO
    gz = golden(f"panoptic_{tag}")
    N, C, T, H, W, seed = (int(gz[k]) for k in "N C T H W seed".split())
    mc, mp, me = synth.panoptic_case(seed, N, C, T, H, W)
    meta = PO.Metadata(*synth.panoptic_metadata(C))
    seg, dic, segments = PO.panoptic_mask_inference(mc.numpy(), mp.numpy(), me.numpy(), meta, pixel_thr=float(gz["thr"]))
    assert np.array_equal(seg, gz["seg"])
    cats = sorted(dic.keys())
    assert cats == list(gz["cats"]) and [len(dic[c]) for c in cats] == list(gz["counts"])
    if cats:
        _close(np.concatenate([np.stack(dic[c]) for c in cats]), gz["embs"], atol=1e-6)
    # every id in the map is -1, a stuff category or category * divisor + instance index of an opened thing segment
    ids = {s[3] for s in segments}
    assert set(np.unique(seg).tolist()) <= ids | {-1}


# --------------------------------------------------------------------------------------------- kMaX axial attention (row f3)
@pytest.mark.parametrize("tag", ["a", "b"])
def test_kmax_axial_attention(golden, tag):
    from oracle import kmax_oracle as KO
    gz = golden(f"kmax_axial_{tag}")
    N, C, L, seed = (int(gz[k]) for k in "N C L seed".split())
    p = synth.kmax_axial_params(seed, C)
    assert synth.checksum(p) == pytest.approx(float(gz["wsum"]), rel=1e-12)
    _close(KO.axial_attention(synth.randn(seed + 100, N, C, L), p), gz["y"])


def test_kmax_axial_attention_2d(golden):
    from oracle import kmax_oracle as KO
    gz = golden("kmax_axial_2d")
    N, C, H, W, seed = (int(gz[k]) for k in "N C H W seed".split())
    ph, pw = synth.kmax_axial_params(seed, C), synth.kmax_axial_params(seed + 1, 1024)
    _close(KO.axial_attention_2d(synth.randn(seed + 100, N, C, H, W), ph, pw), gz["y"])


# ------------------------------------------------------------------------------------------------ Tube-Link (goldens from the TL sources)
def test_tube_link_temporal_encoder(golden):
    """Oracle against the UNMODIFIED Tube-Link TemporalEncoder (source slice exec'd by oracle/make_golden_tl.py): pins rows A1/A2/A7 on TL."""
    gz = golden("tl_temporal")
    B, T, H, W, seed = (int(gz[k]) for k in "B T H W seed".split())
    p = synth.encoder_params(seed, 1)
    assert synth.checksum(p) == pytest.approx(float(gz["wsum"]), rel=1e-12)
    src = synth.randn(seed + 1, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 2)[0])
    y, _, _ = O.temporal_encoder(src, pos, O.split_encoder_params(p))
    assert (y - torch.from_numpy(gz["y"])).abs().max().item() < 2e-5


def test_tube_link_cc_layer(golden):
    """Oracle against the unmodified Tube-Link cross-clip TrajectoryAttentionLayer (TL cc head :152-247)."""
    gz = golden("tl_cc_layer")
    b, Q, T, seed = (int(gz[k]) for k in "b Q T seed".split())
    p = {}
    g = torch.Generator().manual_seed(seed)
    synth.traj_attn_params(g, "self_attn.", 256, p, fused_qkv=True)
    p["norm.weight"] = 1 + 0.1 * torch.randn(256, generator=g)
    p["norm.bias"] = 0.1 * torch.randn(256, generator=g)
    assert synth.checksum(p) == pytest.approx(float(gz["wsum"]), rel=1e-12)
    x = synth.randn(seed + 1, b, T * Q, 256)
    y = O.cc_attention_layer(x, p, Q, T)
    assert (y - torch.from_numpy(gz["y"])).abs().max().item() < 2e-5


def test_tl_decoder_layer_oracle_against_torch_mha(golden):
    """Row A11 (Tube-Link half): the restated decoder layer against the same layer composed from torch.nn.MultiheadAttention / LayerNorm /
    Linear (the modules mmcv's MultiheadAttention / BaseTransformerLayer / FFN wrap; mmcv itself is not installed: wrapper semantics
    unpinned), on the golden inputs and on a second case with an un-masked self-attention only."""
    from axial_vs_b200 import synth
    from oracle import tl_decoder_oracle as TO
    gz = golden("tl_decoder_layer")
    seed, Nq, B, L = (int(gz[k]) for k in "seed Nq B L".split())
    p = synth.tl_decoder_layer_params(seed)
    q, qp, k, kp, m = synth.tl_decoder_case(seed + 1, Nq, B, L)
    y = TO.decoder_layer(q, k, k, qp, kp, [m, None], p)
    assert float((y - torch.from_numpy(gz["y"])).abs().max()) < 2e-5
    q, qp, k, kp, m = synth.tl_decoder_case(seed + 9, 7, 3, 33)
    a = TO.decoder_layer(q, k, k, qp, kp, None, p)
    b = TO.torch_module_composition(q, k, k, qp, kp, None, p)
    assert float((a - b).abs().max()) < 2e-5


@pytest.mark.parametrize("tag", ["a", "b"])
def test_kmax_layer_oracle_golden(golden, tag):
    """Row A11 (Video-kMaX half): the restated kMaXTransformerLayer + kMaXPredictor against the output of the UNMODIFIED reference layer
    (tests/golden/kmax_layer_*.npz from oracle/make_golden_kmax_layer.py; parameters regenerated from the seed)."""
    from oracle import kmax_layer_oracle as KO
    gz = golden(f"kmax_layer_{tag}")
    N, L, Cp, TH, W, K, seed = (int(gz[k]) for k in "N L Cp TH W K seed".split())
    p = synth.kmax_layer_params(seed, Cp, K)
    q, pred = KO.transformer_layer(synth.randn(seed + 100, N, Cp, TH, W), synth.randn(seed + 200, N, 256, L), p)
    for name, got in (("query", q), ("class_logits", pred["class_logits"]), ("mask_logits", pred["mask_logits"]),
                      ("mask_embeddings", pred["mask_embeddings"]), ("pixel_feature", pred["pixel_feature"])):
        want = torch.from_numpy(gz[name])
        assert got.shape == want.shape, name
        assert float((got - want).abs().max() / want.abs().max()) < 2e-5, name
