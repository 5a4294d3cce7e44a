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
