"""CPU: the C-ABI library loads, exports every symbol include/axvs.h declares, validates arguments without a GPU,
and the drop-in modules keep the reference's state-dict keys."""
import ctypes
import os
import re

import pytest
import torch

from axial_vs_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "axvs.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(axvs_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libaxvs.so"


def test_argument_validation_without_gpu(lib):
    assert lib.axvs_version() >= 100
    assert lib.axvs_packed_weight_bytes(768, 256) == 768 * 256 * 2
    assert lib.axvs_pack_weight(None, 8, 64, None, None) == -1
    assert b"null pointer" in lib.axvs_last_error()
    one = ctypes.c_void_p(16)
    assert lib.axvs_pack_weight(one, 7, 64, one, None) == -2           # n_out % 8
    assert lib.axvs_linear(one, 256, 10, 100, one, None, 256, 1.0, 0, one, 256, 1, None, None) == -2   # K % 64
    assert lib.axvs_traj_attn_fwd(one, one, one, None, 0, None, one, None, 1, 2, 3, 4, 1, one, 0, None) == -1
    assert lib.axvs_layer_workspace_bytes(1, 2, 41, 41, 1024) > 0
    assert lib.axvs_layer_workspace_bytes(0, 2, 41, 41, 1024) == 0


def test_workspace_sizes_monotone(lib):
    a = lib.axvs_traj_attn_workspace_bytes(1, 2, 21, 21)
    b = lib.axvs_traj_attn_workspace_bytes(1, 2, 41, 41)
    c = lib.axvs_traj_attn_workspace_bytes(2, 2, 41, 41)
    assert 0 < a < b < c


def test_state_dict_keys_match_reference_names():
    from axial_vs_b200 import modules
    enc = modules.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "axial-trajectory", 2)
    want = synth.encoder_params(0, 2)
    got = enc.state_dict()
    assert set(got) == set(want)
    for k in want:
        assert tuple(got[k].shape) == tuple(want[k].shape), k
    enc2 = modules.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "trajectory", 1)
    assert set(enc2.state_dict()) == set(synth.encoder_params(0, 1, axial=False))
    # reference quirk: the config default "axial_trajectory" (underscore) creates no layers
    assert not hasattr(modules.TemporalEncoder(temporal_attn_type="axial_trajectory"), "temporal_layers")


def test_state_dict_keys_match_live_reference():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not mounted")
    from axial_vs_b200 import modules
    TA = ref_loader.temporal_attention()
    ref = TA.TemporalEncoder(256, 1024, 0.1, 0.1, "relu", 8, "axial-trajectory", 2)
    ours = modules.TemporalEncoder(256, 1024, 0.1, 0.1, "relu", 8, "axial-trajectory", 2)
    assert list(ref.state_dict()) == list(ours.state_dict())
    ours.load_state_dict(ref.state_dict(), strict=True)


def test_cpu_inputs_raise():
    from axial_vs_b200 import modules, ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.pack_weight(torch.zeros(8, 64))
    layer = modules.TemporalAxialTrajectoryAttentionLayer().eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.zeros(2, 4, 256), torch.zeros(1, 2, 2, 2, 256))


def test_unsupported_sizes_raise():
    from axial_vs_b200 import modules
    with pytest.raises(NotImplementedError):
        modules.TrajectoryAttention(128, 4)
    with pytest.raises(NotImplementedError):
        modules.TemporalAxialTrajectoryAttentionLayer(activation="gelu")


def test_kmax_axial_state_dict_keys_match_live_reference():
    """Row f3 drop-ins: same constructor arguments and state-dict keys / shapes as the reference's AxialAttention2D."""
    import sys
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not mounted")
    from axial_vs_b200 import kmax_axial
    ref_loader.cross_clip()                     # loads kmax_pixel_decoder.py
    KP = sys.modules["kmax_deeplab.modeling.pixel_decoder.kmax_pixel_decoder"]
    ref = KP.AxialAttention2D(512, query_shape=[21, 41], filters=512, key_expansion=1, value_expansion=2, num_heads=8)
    ours = kmax_axial.AxialAttention2D(512, query_shape=[21, 41], filters=512, key_expansion=1, value_expansion=2, num_heads=8)
    rs, os_ = ref.state_dict(), ours.state_dict()
    assert list(rs) == list(os_)
    assert all(tuple(rs[k].shape) == tuple(os_[k].shape) for k in rs)
    ours.load_state_dict(rs, strict=True)


def test_panoptic_oracle_invariants():
    """Properties of the mask-wise merge that hold for any input (numpy oracle, CPU): ids come from the segment table, every
    painted pixel had its slot above the pixel threshold, a stuff category appears under one id, thing ids of a category are consecutive."""
    import numpy as np
    from axial_vs_b200 import synth
    from oracle import panoptic_oracle as PO
    for seed in range(5):
        N, C, T, H, W = 24, 9, 2, 17, 13
        mc, mp, me = synth.panoptic_case(100 + seed, N, C, T, H, W, cell=3)
        meta = PO.Metadata(*synth.panoptic_metadata(C, label_divisor=100))
        seg, dic, segments = PO.panoptic_mask_inference(mc.numpy(), mp.numpy(), me.numpy(), meta, pixel_thr=0.3)
        _, _, binary, _, _, _ = PO.scores(mc.numpy(), mp.numpy(), 0.3)
        ids = {s[3] for s in segments}
        assert set(np.unique(seg).tolist()) <= ids | {-1}
        covered = binary.any(0).reshape(seg.shape)
        assert not (seg[~covered] != -1).any()                       # a pixel no slot claims stays unassigned
        stuff_ids = [s[3] for s in segments if not s[2]]
        assert len(stuff_ids) == len(set(stuff_ids))
        by_cat = {}
        for slot, label, is_thing, fid in segments:
            if is_thing:
                by_cat.setdefault(fid // 100, []).append(fid % 100)
        assert all(v == list(range(len(v))) for v in by_cat.values())
        assert {k: len(v) for k, v in dic.items()} == {k: len(v) for k, v in by_cat.items()}


def test_shared_pos_detection():
    """`ops.shared_pos`: a stride-0 broadcast over the clip dim collapses to its single table; materialised tensors pass through."""
    from axial_vs_b200 import ops
    one = torch.randn(1, 2, 3, 4, 8)
    ex = one.expand(5, -1, -1, -1, -1)
    assert ops.shared_pos(ex).shape[0] == 1 and ops.shared_pos(ex).data_ptr() == one.data_ptr()
    full = ex.contiguous()
    assert ops.shared_pos(full) is full
    assert ops.shared_pos(one) is one
    assert ops.shared_pos(ex[1:4]).shape[0] == 1          # slices of a broadcast (clip chunks) stay broadcasts


def test_registry_surface_without_frameworks():
    """`from_config` maps the reference's cfg keys one for one (WC/maxtron_within_clip_tracking_module.py:44-63) and `register()`
    degrades gracefully when detectron2 / mmcv are absent (they are, in this image)."""
    from types import SimpleNamespace as NS
    from axial_vs_b200 import registry
    wc = NS(DROPOUT=0.0, ATTN_DROP=0.0, NHEADS=8, DIM_FEEDFORWARD=1024, NUM_STAGES=2, SPATIAL_LAYERS=2, TEMPORAL_LAYERS=4,
            TEMPORAL_ATTN_TYPE="axial-trajectory", CONV_DIMS=256, SPATIAL_IN_FEATURES=["res3", "res4", "res5"], TEMPORAL_IN_FEATURES=["res4", "res5"])
    cfg = NS(MODEL=NS(MAXTRON=NS(WITHIN_CLIP_TRACKING_MODULE=wc, CROSS_CLIP_TRACKING_MODULE=NS(ENABLE=False))), INPUT=NS(NUM_CLIP_FRAMES=2))
    shape = {f"res{i}": NS(channels=c, stride=s) for i, c, s in ((2, 256, 4), (3, 512, 8), (4, 1024, 16), (5, 2048, 32))}
    kw = registry.B200WithinClipTrackingModule.from_config(cfg, shape)
    assert sorted(kw["input_shape"]) == ["res3", "res4", "res5"] and kw["transformer_temporal_layers"] == 4 and kw["num_clip_frames"] == 2
    m = registry.B200WithinClipTrackingModule(**kw)
    keys = list(m.state_dict().keys())
    assert all(k.startswith("within_clip_tracking_module.") for k in keys)                    # the reference's checkpoint prefix
    assert any("transformer.encoder.temporal_layers.1.temporal_layers.1.width_attn.proj_kv.weight" in k for k in keys)
    out = registry.register()
    assert set(out) == {"detectron2", "mmcv"}
