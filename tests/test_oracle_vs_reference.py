"""CPU, authoring container only: oracle vs the live, unmodified reference modules (skips when
/root/reference is absent, e.g. on the GPU box)."""
import pytest
import torch

from axial_vs_b200 import synth
from oracle import ref_loader
from oracle import traj_oracle as O

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


@torch.no_grad()
def test_axial_encoder_live():
    TA = ref_loader.temporal_attention()
    B, T, H, W, seed = 2, 2, 7, 9, 77
    p = synth.encoder_params(seed, 2)
    m = TA.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "axial-trajectory", 2).eval()
    m.load_state_dict(p, strict=True)
    src = synth.randn(seed + 1, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 2)[0])
    ref, hm, wm = m(src, pos)
    out, ohm, owm = O.temporal_encoder(src, pos, O.split_encoder_params(p), return_maps=True)
    assert (ref - out).abs().max().item() < 3e-5
    assert (hm - ohm).abs().max().item() < 1e-5 and (wm - owm).abs().max().item() < 1e-5


@torch.no_grad()
def test_pos3d_live():
    PE = ref_loader.pos_embeddings()
    pe = PE.PositionEmbeddingSine3D(128, normalize=True)
    tab = pe(torch.zeros(2, 3, 256, 11, 13)).permute(0, 1, 3, 4, 2)
    assert (tab - O.pos3d_table(2, 3, 11, 13)).abs().max().item() < 1e-6


@torch.no_grad()
def test_cross_clip_live():
    CC = ref_loader.cross_clip()
    Q, Tc, V, Hh, Ww, L, K, seed = 24, 5, 2, 4, 6, 2, 124, 88
    p = synth.cross_clip_params(seed, L, K)
    m = CC.CrossClipTrackingModule(num_layers=L, num_classes=K, attn_drop=0.0, aspp_drop=0.0,
                                   kernel_sizes=[3, 3, 3], atrous_rates=[1, 2, 3], norm_fn="ln",
                                   num_clip_frames=V).eval()
    m.load_state_dict(p, strict=True)
    cq = synth.randn(seed + 1, 1, Q, Tc, 256)
    pf = synth.randn(seed + 2, 1, 128, Tc * V, Hh, Ww)
    ref = m(cq, pf)
    o = O.cross_clip_module(cq, pf, p, L, V)
    assert (ref["pred_logits"] - o["pred_logits"]).abs().max().item() < 3e-5
    assert (ref["pred_masks"] - o["pred_masks"]).abs().max().item() < 1e-4
