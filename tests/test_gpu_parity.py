"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed reference goldens.

Tolerance (BASELINE.json north_star): bf16 compute / fp32 accumulate, max-abs error normalised by the
max-abs of the reference tensor <= 1e-2 on attention / layer outputs.
"""
import numpy as np
import pytest
import torch

from axial_vs_b200 import synth

pytestmark = pytest.mark.gpu

TOL = 1e-2


def nerr(a: torch.Tensor, ref: torch.Tensor) -> float:
    a, ref = a.detach().float().cpu(), ref.detach().float().cpu()
    assert a.shape == ref.shape, (a.shape, ref.shape)
    assert torch.isfinite(a).all(), "non-finite values in CUDA output"
    return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def ops():
    from axial_vs_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def O():
    from oracle import traj_oracle
    return traj_oracle


# --------------------------------------------------------------------------------------------- GEMM engine
@pytest.mark.parametrize("M,K,N", [(128, 256, 256), (300, 256, 512), (1000, 1024, 256), (77, 64, 768), (20000, 256, 1024)])
def test_linear_bf16(ops, M, K, N):
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    out = ops.linear(a, ops.pack_weight(w), b, N)
    ref = a.float() @ w.bfloat16().float().t() + b
    assert nerr(out, ref) < 6e-3   # bf16 output rounding only


def test_linear_epilogues(ops):
    g = torch.Generator().manual_seed(7)
    M, K, N = 513, 256, 256
    a = torch.randn(M, K, generator=g).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) / 16).cuda()
    b = torch.randn(N, generator=g).cuda()
    r = torch.randn(M, N, generator=g).cuda()
    wp = ops.pack_weight(w)
    base = a.float() @ w.bfloat16().float().t() + b
    assert nerr(ops.linear(a, wp, b, N, out_dtype=torch.float32), base) < 1e-5
    assert nerr(ops.linear(a, wp, b, N, out_dtype=torch.float32, resid=r), base + r) < 1e-5
    assert nerr(ops.linear(a, wp, b, N, relu=True, out_dtype=torch.float32), base.relu()) < 1e-5
    assert nerr(ops.linear(a, wp, b, N, scale=0.25, out_dtype=torch.float32), base * 0.25) < 1e-5
    assert nerr(ops.linear(a, wp, None, N, out_dtype=torch.float32), base - b) < 1e-5


# --------------------------------------------------------------------------------------------- spatial attention
@pytest.mark.parametrize("S,F,n", [(3, 2, 5), (2, 3, 41), (1, 2, 70), (2, 1, 130), (1, 5, 64)])
def test_spatial_attention(ops, S, F, n):
    g = torch.Generator().manual_seed(S * 100 + F * 10 + n)
    N = F * n
    qkv = torch.randn(S * N, 768, generator=g).bfloat16().cuda()
    x = ops.spatial_attention(qkv, S, F, n)                      # [S*N, F, 256]
    q, k, v = (t.float().reshape(S, N, 8, 32).permute(0, 2, 1, 3) for t in qkv.split(256, dim=1))
    ref = torch.empty(S, N, F, 256, device="cuda")
    for f in range(F):
        kf, vf = k[:, :, f * n:(f + 1) * n], v[:, :, f * n:(f + 1) * n]
        a = torch.softmax(32 ** -0.5 * q @ kf.transpose(-1, -2), -1)
        ref[:, :, f] = (a @ vf).permute(0, 2, 1, 3).reshape(S, N, 256)
    assert nerr(x.reshape(S, N, F, 256), ref) < 1e-2


# --------------------------------------------------------------------------------------------- trajectory attention
def _ta_case(ops, O, Bp, F, n, seed, fused_qkv=False):
    p = {}
    synth.traj_attn_params(torch.Generator().manual_seed(seed), "", 256, p, fused_qkv=fused_qkv)
    q = synth.randn(seed + 100, Bp, F * n, 256)
    v = synth.randn(seed + 200, Bp, F * n, 256)
    pk = ops.pack_ta({k_: t.cuda() for k_, t in p.items()})
    return p, q, v, pk


@pytest.mark.parametrize("tag", ["a", "b"])
def test_trajectory_attention_golden(ops, golden, tag):
    gz = golden(f"ta_vk_{tag}")
    Bp, F, n, seed = int(gz["Bp"]), int(gz["F"]), int(gz["n"]), int(gz["seed"])
    p, q, v, pk = _ta_case(ops, None, Bp, F, n, seed)
    assert synth.checksum(p) == pytest.approx(float(gz["wsum"]), rel=1e-12)
    qc, vc = q.reshape(-1, 256).cuda(), v.reshape(-1, 256).cuda()
    out = ops.traj_attn_fwd(qc, qc, vc, None, None, pk, Bp, F, n, 1, ops.AXIS_NONE)
    assert nerr(out.reshape(Bp, F * n, 256), torch.from_numpy(gz["y"])) < TOL


@pytest.mark.parametrize("Bp,F,n", [(5, 2, 41), (2, 5, 30), (1, 10, 33), (1, 2, 200), (1, 2, 161), (2, 3, 100), (3, 2, 21), (1, 4, 64), (1, 2, 128),
                                    (3, 1, 40), (70, 3, 1), (40, 2, 3), (300, 2, 7)])   # one frame; one token per frame; short runs over many tiles
def test_trajectory_attention_oracle(ops, O, Bp, F, n):
    p, q, v, pk = _ta_case(ops, O, Bp, F, n, 1000 + Bp + F + n)
    ref, _ = O.trajectory_attention(q, q, v, p, F)
    qc, vc = q.reshape(-1, 256).cuda(), v.reshape(-1, 256).cuda()
    out = ops.traj_attn_fwd(qc, qc, vc, None, None, pk, Bp, F, n, 1, ops.AXIS_NONE)
    assert nerr(out.reshape(Bp, F * n, 256), ref) < TOL


def test_trajectory_attention_module_separate_key(O):
    """nn.Module API with key is not query (general reference signature)."""
    from axial_vs_b200 import modules
    p = {}
    synth.traj_attn_params(torch.Generator().manual_seed(9), "", 256, p)
    q, k, v = synth.randn(1, 2, 24, 256), synth.randn(2, 2, 24, 256), synth.randn(3, 2, 24, 256)
    ref, _ = O.trajectory_attention(q, k, v, p, 3)
    m = modules.TrajectoryAttention(256, 8, 0.0).eval()
    m.load_state_dict(p, strict=True)
    m.cuda()
    with torch.no_grad():
        out, maps = m(q.cuda(), k.cuda(), v.cuda(), num_frames=3)
    assert maps is None
    assert nerr(out, ref) < TOL


# --------------------------------------------------------------------------------------------- layers / encoder
def _layer(p, axial=True):
    from axial_vs_b200 import modules
    cls = modules.TemporalAxialTrajectoryAttentionLayer if axial else modules.TemporalTrajectoryAttentionLayer
    m = cls(256, 1024, 0.0, 0.0, "relu", 8).eval()
    m.load_state_dict(p, strict=True)
    return m.cuda()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_axial_layer_golden(O, golden, tag):
    gz = golden(f"axial_layer_{tag}")
    B, T, H, W, seed = (int(gz[k]) for k in "B T H W seed".split())
    p = synth.axial_layer_params(seed)
    src = synth.randn(seed + 100, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 200)[0])
    with torch.no_grad():
        out, hm, wm = _layer(p)(src.cuda(), pos.cuda())
    assert hm is None and wm is None
    assert nerr(out, torch.from_numpy(gz["out"])) < TOL


def test_encoder_golden(O, golden):
    from axial_vs_b200 import modules
    gz = golden("encoder_axial")
    B, T, H, W, seed = (int(gz[k]) for k in "B T H W seed".split())
    p = synth.encoder_params(seed, 2)
    enc = modules.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "axial-trajectory", 2).eval()
    enc.load_state_dict(p, strict=True)
    enc.cuda()
    src = synth.randn(seed + 100, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 200)[1])
    with torch.no_grad():
        out, _, _ = enc(src.cuda(), pos.cuda())
    assert nerr(out, torch.from_numpy(gz["out"])) < TOL


def test_trajectory_encoder_golden(O, golden):
    from axial_vs_b200 import modules
    gz = golden("encoder_trajectory")
    B, T, H, W, seed = (int(gz[k]) for k in "B T H W seed".split())
    p = synth.encoder_params(seed, 1, axial=False)
    enc = modules.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "trajectory", 1).eval()
    enc.load_state_dict(p, strict=True)
    enc.cuda()
    src = synth.randn(seed + 100, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 200)[0])
    with torch.no_grad():
        out, _, _ = enc(src.cuda(), pos.cuda())
    assert nerr(out, torch.from_numpy(gz["out"])) < TOL


@pytest.mark.parametrize("B,T,H,W", [(1, 2, 41, 41), (1, 2, 21, 21), (2, 5, 15, 20), (3, 2, 13, 29)])
def test_axial_layer_oracle_config_sizes(O, B, T, H, W):
    """BASELINE config 1 (T=2, 41x41), res5 level, Tube-Link-shaped (T=5, 15x20) and a ragged batch."""
    seed = 500 + B + T + H + W
    p = synth.axial_layer_params(seed)
    src = synth.randn(seed + 1, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 2)[0])
    ref, _, _ = O.axial_layer(src, pos, p)
    with torch.no_grad():
        out, _, _ = _layer(p)(src.cuda(), pos.cuda())
    e = nerr(out, ref)
    cos = torch.nn.functional.cosine_similarity(out.cpu().flatten(), ref.flatten(), dim=0).item()
    assert e < TOL and cos > 0.9999, (e, cos)


@pytest.mark.parametrize("B,T,H,W", [(2, 2, 41, 41), (1, 5, 15, 20), (3, 2, 13, 29), (5, 2, 21, 21), (2, 3, 7, 40), (1, 1, 50, 3)])
def test_direct_qkv_matches_packed_front_end(ops, B, T, H, W):
    """Fusion level 4 (q|k|v GEMM reading the fp32 stream, every A operand in tensor memory, q2 held as bf16) against
    level 3 (tile-image pack kernel, shared-memory operands, fp32 q2): same bf16 GEMM operands otherwise."""
    seed = 900 + B + T + H + W
    p = synth.axial_layer_params(seed)
    src = synth.randn(seed + 1, B * T, H * W, 256).cuda()
    pos = synth.randn(seed + 2, B, T, H, W, 256).cuda()
    layer = _layer(p)
    outs = []
    try:
        for level, core in ((3, 1), (4, 1), (4, 0)):
            ops.set_fusion(level)
            ops.set_attn_core(core)
            with torch.no_grad():
                outs.append(layer(src, pos)[0])
            torch.cuda.synchronize()
    finally:
        ops.set_fusion(ops.DEFAULT_FUSION)
        ops.set_attn_core(1)
    assert torch.isfinite(outs[1]).all()
    assert nerr(outs[1], outs[0]) < 4e-3
    assert nerr(outs[2], outs[1]) < 4e-3      # mma.sync attention core against the tcgen05 core


@pytest.mark.parametrize("B,T,H,W", [(3, 2, 13, 29), (5, 2, 21, 21), (2, 3, 7, 40)])
def test_shared_pos_table_is_bit_identical(ops, B, T, H, W):
    """One [1,T,H,W,C] positional table shared by the clips (or its stride-0 expand, what PositionEmbeddingSine3D.table returns)
    must give exactly the result of the materialised [B,T,H,W,C] tensor, at every fusion level."""
    seed = 1300 + B + T + H + W
    p = synth.axial_layer_params(seed)
    src = synth.randn(seed + 1, B * T, H * W, 256).cuda()
    one = synth.randn(seed + 2, 1, T, H, W, 256).cuda()
    full = one.expand(B, -1, -1, -1, -1).contiguous()
    layer = _layer(p)
    try:
        for level in (0, 2, 3, 4):
            ops.set_fusion(level)
            with torch.no_grad():
                ref = layer(src, full)[0]
                a = layer(src, one)[0]
                b = layer(src, one.expand(B, -1, -1, -1, -1))[0]
            torch.cuda.synchronize()
            assert torch.equal(a, ref) and torch.equal(b, ref), f"fusion level {level}"
    finally:
        ops.set_fusion(ops.DEFAULT_FUSION)
    with pytest.raises(RuntimeError):
        layer(src, full[:2].contiguous()) if B > 2 else layer(src, torch.cat([full, full]).contiguous())


def test_encoder_config1_two_layers(O):
    """BASELINE config 1: TemporalEncoder(axial-trajectory, 2 layers) on T=2, 41x41x256."""
    from axial_vs_b200 import modules
    B, T, H, W, seed = 1, 2, 41, 41, 0
    p = synth.encoder_params(seed, 2)
    enc = modules.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "axial-trajectory", 2).eval()
    enc.load_state_dict(p, strict=True)
    enc.cuda()
    src = synth.randn(1, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(2)[1])
    ref, _, _ = O.temporal_encoder(src, pos, O.split_encoder_params(p))
    with torch.no_grad():
        out, _, _ = enc(src.cuda(), pos.cuda())
    assert nerr(out, ref) < TOL


# --------------------------------------------------------------------------------------------- positional table
@pytest.mark.parametrize("tag", ["a", "b"])
def test_pos3d_golden(ops, golden, tag):
    gz = golden(f"pos3d_{tag}")
    B, T, H, W = (int(gz[k]) for k in "B T H W".split())
    out = ops.pos3d(B, T, H, W, None, "cuda")
    assert (out.cpu() - torch.from_numpy(gz["table"])).abs().max().item() < 2e-5
    le = synth.level_embed(3)[1].cuda()
    out2 = ops.pos3d(B, T, H, W, le, "cuda")
    assert (out2 - out - le).abs().max().item() < 1e-6


# --------------------------------------------------------------------------------------------- properties / errors
def test_batch_independence(O):
    """Clips are independent (the sharding invariant): batching B clips == running them one at a time."""
    B, T, H, W = 3, 2, 9, 11
    p = synth.axial_layer_params(123)
    layer = _layer(p)
    src = synth.randn(5, B * T, H * W, 256).cuda()
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(6)[0]).cuda()
    with torch.no_grad():
        full, _, _ = layer(src, pos)
        for b in range(B):
            one, _, _ = layer(src[b * T:(b + 1) * T].contiguous(), pos[b:b + 1].contiguous())
            assert torch.equal(one, full[b * T:(b + 1) * T]), "per-clip result differs from the batched result"


def test_cpu_tensor_is_rejected():
    from axial_vs_b200 import modules
    layer = modules.TemporalAxialTrajectoryAttentionLayer().eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.zeros(2, 4, 256), torch.zeros(1, 2, 2, 2, 256))


def test_sweep_max_size_sequence_independence():
    """BASELINE configs[4] corner (T=10, 161x161: 161 sequences of 1610 tokens, frames of 161 keys) -- far beyond what the CPU
    oracle can check, so use the size-independent property: every sequence's result is independent of the rest of the batch
    (bit-exact), finite, and the attention over a single frame (F=1 slice) matches the oracle on a small sub-problem."""
    from axial_vs_b200.modules import TrajectoryAttention
    g = torch.Generator().manual_seed(4242)
    p = {}
    synth.traj_attn_params(g, "", 256, p)
    ta = TrajectoryAttention(256, 8, 0.0).eval()
    ta.load_state_dict(p, strict=True)
    ta.cuda()
    Bp, F, n = 161, 10, 161
    x = torch.randn(Bp, F * n, 256, generator=g).cuda()
    with torch.no_grad():
        full, _ = ta(x, x, x, num_frames=F)
        part, _ = ta(x[:3].contiguous(), x[:3].contiguous(), x[:3].contiguous(), num_frames=F)
    assert torch.isfinite(full).all()
    assert torch.equal(full[:3], part)


def test_empty_batch_returns_empty():
    """Zero clips: the modules return empty tensors of the right shape without launching anything (reference behaviour)."""
    p = synth.axial_layer_params(5)
    layer = _layer(p)
    with torch.no_grad():
        out, hm, wm = layer(torch.empty(0, 6 * 5, 256, device="cuda"), torch.empty(0, 2, 6, 5, 256, device="cuda"))
    assert out.shape == (0, 30, 256) and hm is None and wm is None
    from axial_vs_b200.modules import TrajectoryAttention
    ta = TrajectoryAttention(256, 8, 0.0).eval().cuda()
    with torch.no_grad():
        y, maps = ta(torch.empty(0, 10, 256, device="cuda"), torch.empty(0, 10, 256, device="cuda"), torch.empty(0, 10, 256, device="cuda"), num_frames=2)
    assert y.shape == (0, 10, 256) and maps is None


def test_training_mode_is_rejected():
    from axial_vs_b200 import modules
    layer = modules.TemporalAxialTrajectoryAttentionLayer().cuda().train()
    with pytest.raises(RuntimeError, match="inference"):
        layer(torch.zeros(2, 4, 256, device="cuda"), torch.zeros(1, 2, 2, 2, 256, device="cuda"))


# --------------------------------------------------------------------------------------------- fusion levels
@pytest.mark.parametrize("Bp,F,n", [(5, 2, 41), (3, 5, 30), (1, 10, 33), (7, 2, 21), (1, 1, 50)])
def test_fusion_levels_agree(ops, O, Bp, F, n):
    """Every fusion level of the composite call must match the oracle, and the levels must agree with each other."""
    p, q, v, pk = _ta_case(ops, O, Bp, F, n, 2000 + Bp + F + n)
    ref, _ = O.trajectory_attention(q, q, v, p, F)
    qc, vc = q.reshape(-1, 256).cuda(), v.reshape(-1, 256).cuda()
    res = torch.randn(Bp * F * n, 256, generator=torch.Generator().manual_seed(1)).cuda()
    outs = []
    try:
        for level in (0, 1, 2, 3, 4):
            ops.set_fusion(level)
            out = ops.traj_attn_fwd(qc, qc, vc, None, res, pk, Bp, F, n, 1, ops.AXIS_NONE)
            torch.cuda.synchronize()
            assert nerr(out.cpu() - res.cpu(), ref.reshape(-1, 256)) < TOL, f"fusion level {level}"
            outs.append(out)
    finally:
        ops.set_fusion(ops.DEFAULT_FUSION)
    assert nerr(outs[1], outs[0]) < 5e-3 and nerr(outs[3], outs[0]) < 5e-3
    assert nerr(outs[4], outs[3]) < 4e-3      # level 4: tensor-memory operands, q2 rounded to bf16


@pytest.mark.parametrize("Bp,F,n", [(5, 2, 41), (7, 2, 21), (3, 5, 30), (2, 5, 40), (1, 10, 33), (2, 2, 5), (1, 1, 50), (1, 2, 161), (1, 2, 200),
                                    (1, 3, 224), (2, 2, 128), (1, 4, 64), (3, 1, 16), (1, 7, 17), (9, 2, 81)])
def test_attention_core_tcgen05(ops, O, Bp, F, n):
    """tcgen05 attention core (S = Q K^T and P V as UMMAs, softmax out of tensor memory) against the mma.sync core and
    the oracle; query == key == value input so that the level-4 front end (qkv_direct) runs, sequences longer than one
    128-query block included."""
    p, q, _, pk = _ta_case(ops, O, Bp, F, n, 3000 + Bp + F + n)
    ref, _ = O.trajectory_attention(q, q, q, p, F)
    qc = q.reshape(-1, 256).cuda()
    outs = []
    try:
        for core in (0, 1):
            ops.set_attn_core(core)
            out = ops.traj_attn_fwd(qc, qc, qc, None, None, pk, Bp, F, n, 1, ops.AXIS_NONE)
            torch.cuda.synchronize()
            assert nerr(out, ref.reshape(-1, 256)) < TOL, f"attention core {core}"
            outs.append(out)
    finally:
        ops.set_attn_core(1)
    assert nerr(outs[1], outs[0]) < 4e-3


# --------------------------------------------------------------------------------------------- maps, pos module, TL, cross-clip
def test_attention_maps_golden(O, golden):
    """return_attn_maps=True reproduces the reference's `space_attn` tuple members (layout [(B' h), N, F, n], fp32)."""
    gz = golden("axial_layer_a")
    B, T, H, W, seed = (int(gz[k]) for k in "B T H W seed".split())
    p = synth.axial_layer_params(seed)
    layer = _layer(p)
    layer.return_attn_maps = True
    src = synth.randn(seed + 100, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 200)[0])
    with torch.no_grad():
        out, hm, wm = layer(src.cuda(), pos.cuda())
    assert nerr(out, torch.from_numpy(gz["out"])) < TOL
    assert tuple(hm.shape) == gz["hmap"].shape and tuple(wm.shape) == gz["wmap"].shape
    assert (hm.cpu() - torch.from_numpy(gz["hmap"])).abs().max().item() < 2e-2      # probabilities, absolute tolerance
    assert (wm.cpu() - torch.from_numpy(gz["wmap"])).abs().max().item() < 2e-2
    assert (hm.sum(-1) - 1).abs().max().item() < 1e-4


def test_pos_module_matches_reference_layout(golden):
    from axial_vs_b200.pos import PositionEmbeddingSine3D
    gz = golden("pos3d_b")
    B, T, H, W = (int(gz[k]) for k in "B T H W".split())
    pe = PositionEmbeddingSine3D(128, normalize=True)
    out = pe(torch.zeros(B, T, 256, H, W, device="cuda"), fmt="btchw")
    assert tuple(out.shape) == (B, T, 256, H, W)
    assert (out.permute(0, 1, 3, 4, 2).cpu() - torch.from_numpy(gz["table"])).abs().max().item() < 2e-5


def test_tube_link_wrappers(O):
    """TL flavour: features only, gamma skip (TL msdeformattn_pixel_decoder.py:620-627)."""
    from axial_vs_b200 import tube_link
    B, T, H, W, seed = 1, 5, 6, 8, 321
    p = synth.encoder_params(seed, 1)
    enc = tube_link.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, 1).eval()
    enc.load_state_dict(p, strict=True)
    enc.cuda()
    f = synth.randn(seed + 1, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 2)[0])
    gamma = torch.full((256,), 0.5)
    ref, _, _ = O.temporal_encoder(f, pos, O.split_encoder_params(p))
    ref = f + gamma * ref
    with torch.no_grad():
        out = tube_link.temporal_branch([f.cuda(), f.cuda()], [pos.cuda()], enc, gamma.cuda(), 1)
    assert nerr(out[0], ref) < TOL and torch.equal(out[1].cpu(), f)


def test_cross_clip_attention_golden(golden):
    from axial_vs_b200 import cross_clip
    gz = golden("cc_ta")
    b, Q, T, seed = (int(gz[k]) for k in "b Q T seed".split())
    p = {}
    synth.traj_attn_params(torch.Generator().manual_seed(seed), "", 256, p, fused_qkv=True)
    m = cross_clip.TrajectoryAttention(256, 8, 0.0).eval()
    m.load_state_dict(p, strict=True)
    m.cuda()
    x = synth.randn(seed + 100, b, T * Q, 256)
    with torch.no_grad():
        y = m(x.cuda(), seq_len=Q, num_frames=T)
    assert nerr(y, torch.from_numpy(gz["y"])) < TOL


def test_cross_clip_layer_oracle(O):
    """cfg3-shaped: Q = 128 queries per clip, 8 clips (a per-GPU shard of the 64-clip video)."""
    from axial_vs_b200 import cross_clip
    b, Q, T, seed = 1, 128, 8, 654
    p = {}
    g = torch.Generator().manual_seed(seed)
    synth.traj_attn_params(g, "self_attn.", 256, p, fused_qkv=True)
    p["norm.weight"] = 1 + 0.1 * torch.randn(256, generator=g)
    p["norm.bias"] = 0.1 * torch.randn(256, generator=g)
    m = cross_clip.TrajectoryAttentionLayer(256, 8).eval()
    m.load_state_dict(p, strict=True)
    m.cuda()
    x = synth.randn(seed + 1, b, T * Q, 256)
    ref = O.cc_attention_layer(x, p, Q, T)
    with torch.no_grad():
        y = m(x.cuda(), Q, T)
    assert nerr(y, ref) < TOL


def test_cross_clip_layer_oracle_64_clips(O):
    """BASELINE configs[2] at its FULL length: one video of 64 clips x 128 queries = 8192 tokens, F = 64 frames of n = 128 tokens
    (the tcgen05 attention core with 64 query blocks per sequence and one frame per unit; x [8192, 64, 256] never leaves bf16 images)."""
    from axial_vs_b200 import cross_clip
    b, Q, T, seed = 1, 128, 64, 655
    p = {}
    g = torch.Generator().manual_seed(seed)
    synth.traj_attn_params(g, "self_attn.", 256, p, fused_qkv=True)
    p["norm.weight"] = 1 + 0.1 * torch.randn(256, generator=g)
    p["norm.bias"] = 0.1 * torch.randn(256, generator=g)
    m = cross_clip.TrajectoryAttentionLayer(256, 8).eval()
    m.load_state_dict(p, strict=True)
    m.cuda()
    x = synth.randn(seed + 1, b, T * Q, 256)
    ref = O.cc_attention_layer(x, p, Q, T)
    with torch.no_grad():
        y = m(x.cuda(), Q, T)
    assert nerr(y, ref) < TOL


@pytest.mark.parametrize("B,T,H,W", [(1, 5, 30, 40), (1, 2, 81, 81), (2, 2, 25, 43), (1, 2, 49, 85)])
def test_axial_layer_oracle_large_shapes(O, B, T, H, W):
    """BASELINE configs[3] at its larger temporal level (T = 5, 30x40: sequences of 150 / 200 tokens, two query blocks each), the
    sweep's 81x81 point (sequences of 162 tokens) and the res5 / res4 maps of the shipped VIPSeg configs (769 x 1345 -> 25 x 43 and 49 x 85,
    Vk/configs/VIPSeg/panoptic_segmentation/maxtron_wc_*.yaml: IMAGE_SIZE), whole layer against the oracle."""
    seed = 5000 + T + H + W
    p = synth.axial_layer_params(seed)
    src = synth.randn(seed + 1, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 2)[0])
    ref, _, _ = O.axial_layer(src, pos, p)
    with torch.no_grad():
        out, _, _ = _layer(p)(src.cuda(), pos.cuda())
    assert nerr(out, ref) < TOL


def test_level_plumbing(O):
    from axial_vs_b200 import modules, within_clip
    from axial_vs_b200.pos import PositionEmbeddingSine3D
    B, T, seed = 1, 2, 77
    shapes = [(5, 6), (9, 11), (17, 21)]
    p = synth.encoder_params(seed, 1)
    enc = modules.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "axial-trajectory", 1).eval()
    enc.load_state_dict(p, strict=True)
    enc.cuda()
    le = synth.level_embed(seed + 1)
    mem = synth.randn(seed + 2, B * T, sum(h * w for h, w in shapes), 256)
    pe = PositionEmbeddingSine3D(128, normalize=True)
    pos = within_clip.level_positions(pe, le.cuda(), B, T, shapes[:2], "cuda")
    with torch.no_grad():
        out, _, _ = within_clip.run_temporal_levels(enc, mem.cuda(), shapes, pos, 2)
    parts = list(torch.split(mem, [h * w for h, w in shapes], dim=1))
    for i in range(2):
        parts[i], _, _ = O.temporal_encoder(parts[i].contiguous(), O.level_pos3d(B, T, *shapes[i], le[i]), O.split_encoder_params(p))
    assert nerr(out, torch.cat(parts, 1)) < TOL


def test_cross_clip_module_golden(golden):
    """Full CrossClipTrackingModule (trajectory attention + ASPP + projections + predictor) against the reference output."""
    from axial_vs_b200 import cross_clip
    gz = golden("cc_module")
    Q, T, V, H, W, L, K, seed = (int(gz[k]) for k in "Q T V H W L K seed".split())
    p = synth.cross_clip_params(seed, L, K)
    m = cross_clip.CrossClipTrackingModule(num_layers=L, num_classes=K, attn_drop=0.0, aspp_drop=0.0, kernel_sizes=[3, 3, 3],
                                           atrous_rates=[1, 2, 3], norm_fn="ln", num_clip_frames=V).eval()
    m.load_state_dict(p, strict=True)
    m.cuda()
    cq = synth.randn(seed + 100, 1, Q, T, 256)
    pf = synth.randn(seed + 200, 1, 128, T * V, H, W)
    with torch.no_grad():
        o = m(cq.cuda(), pf.cuda())
    ref_logits, ref_masks = torch.from_numpy(gz["pred_logits"]), torch.from_numpy(gz["pred_masks"])
    assert tuple(o["pred_logits"].shape) == tuple(ref_logits.shape) and tuple(o["pred_masks"].shape) == tuple(ref_masks.shape)
    assert nerr(o["pred_logits"], ref_logits) < TOL
    assert nerr(o["pred_masks"], ref_masks) < TOL           # the mask branch runs split-precision (fp32-grade) after the bf16 attention
    agree = (o["pred_masks"].cpu().argmax(1) == ref_masks.argmax(1)).float().mean().item()
    print(f"[cc_module golden] per-pixel argmax agreement {agree:.5f}")
    assert agree >= 0.995, agree                           # tiny random-init logits: near-ties dominate the disagreements
    assert nerr(o["aux_outputs"][0]["pred_masks"], torch.from_numpy(gz["aux_masks"])) < TOL


def test_cross_clip_module_oracle_cfg3_shard(O):
    """cfg3-shaped shard: Q = 128 queries, 8 clips of 2 frames, 4 layers, 40x40 mask features."""
    from axial_vs_b200 import cross_clip
    Q, T, V, H, W, L, K, seed = 128, 8, 2, 40, 40, 4, 124, 909
    p = synth.cross_clip_params(seed, L, K)
    m = cross_clip.CrossClipTrackingModule(num_layers=L, num_classes=K, attn_drop=0.0, aspp_drop=0.0, kernel_sizes=[3, 3, 3],
                                           atrous_rates=[1, 2, 3], norm_fn="ln", num_clip_frames=V).eval()
    m.load_state_dict(p, strict=True)
    m.cuda()
    cq = synth.randn(seed + 1, 1, Q, T, 256)
    pf = synth.randn(seed + 2, 1, 128, T * V, H, W)
    ref = O.cross_clip_module(cq, pf, p, L, V)
    with torch.no_grad():
        o = m(cq.cuda(), pf.cuda())
    assert nerr(o["pred_logits"], ref["pred_logits"]) < TOL
    assert nerr(o["pred_masks"], ref["pred_masks"]) < TOL
    assert nerr(m.last_clip_query, ref["clip_query"]) < TOL
    # north_star: per-pixel agreement of the final argmax labels (query index per pixel, class per query).  With RANDOM-INIT heads
    # the 128 queries' logits of a pixel are nearly tied (median top-2 margin ~ 4 % of the logit range), so the
    # agreement measures operand rounding directly.  The stages after the bf16 trajectory attention run split-precision
    # (tests/test_error_budget_cpu.py shows each of them cost more agreement in bf16 than the whole attention).
    got, want = o["pred_masks"].float().cpu(), ref["pred_masks"].float()
    agree = (got.argmax(1) == want.argmax(1)).float().mean().item()
    print(f"[cfg3 shard] per-pixel argmax agreement {agree:.5f}, logit error {nerr(got, want):.2e}")
    assert agree >= 0.998, f"per-pixel mask argmax agreement {agree:.5f}"     # measured 0.9988 = the bf16-attention bound of the error budget
    top2 = want.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    err = (got - want).abs().max().item()
    decided = margin > 2.0 * err
    assert decided.float().mean().item() > 0.5
    assert torch.equal(got.argmax(1)[decided], want.argmax(1)[decided])
    assert torch.equal(o["pred_logits"].argmax(-1).cpu(), ref["pred_logits"].argmax(-1))


def test_cross_clip_forward_sharded_equals_unsharded():
    """Clip-sharded cross-clip module: two emulated ranks (5 clips -> 3 + 2) with an injected gather must give
    the unsharded module's final predictions bit for bit (the layers run redundantly, the mask einsum is per clip)."""
    from axial_vs_b200 import cross_clip, sharding
    Q, T, V, H, W, L, K, seed = 32, 5, 2, 12, 10, 2, 19, 911
    p = synth.cross_clip_params(seed, L, K)
    m = cross_clip.CrossClipTrackingModule(num_layers=L, num_classes=K, attn_drop=0.0, aspp_drop=0.0, kernel_sizes=[3, 3, 3],
                                           atrous_rates=[1, 2, 3], norm_fn="ln", num_clip_frames=V).eval()
    m.load_state_dict(p, strict=True)
    m.cuda()
    cq = synth.randn(seed + 1, 1, Q, T, 256).cuda()
    pf = synth.randn(seed + 2, 1, 128, T * V, H, W).cuda()
    with torch.no_grad():
        full = m(cq, pf)
        # what an all-gather would return on every rank: the concatenation of the shards in clip order
        bounds = [sharding.shard_range(T, r, 2) for r in range(2)]
        cq_parts = [cq[0, :, a:b].permute(1, 0, 2).contiguous() for a, b in bounds]
        masks = []
        for rank, (a, b) in enumerate(bounds):
            def gather(x, n, rank=rank):
                if x.shape[1:] == cq_parts[0].shape[1:]:                       # the clip queries
                    return torch.cat(cq_parts, 0)
                masks.append(x)                                                # this rank's mask logits [T_local, Q, P]
                return torch.cat([x.new_zeros((a,) + x.shape[1:]), x, x.new_zeros((T - b,) + x.shape[1:])], 0)
            out = _forward_sharded_as_rank(m, cq[:, :, a:b].contiguous(), pf[:, :, a * V:b * V].contiguous(), T, rank, 2, gather)
            assert torch.equal(out["pred_logits"], full["pred_logits"])
            assert torch.equal(out["pred_masks"][:, :, a * V:b * V], full["pred_masks"][:, :, a * V:b * V])


def _forward_sharded_as_rank(m, cq_local, pf_local, n_clips, rank, world, gather):
    """forward_sharded with shard_range evaluated for an emulated (rank, world) -- no process group in the GPU tests."""
    from axial_vs_b200 import sharding
    import torch.distributed as dist
    orig_ws, orig_rk, orig_init = dist.get_world_size, dist.get_rank, dist.is_initialized
    dist.get_world_size, dist.get_rank, dist.is_initialized = (lambda group=None: world), (lambda group=None: rank), (lambda: True)
    try:
        return m.forward_sharded(cq_local, pf_local, n_clips, gather=gather)
    finally:
        dist.get_world_size, dist.get_rank, dist.is_initialized = orig_ws, orig_rk, orig_init


PAIR_DEFAULT = 30      # traj_pair | qkv_pair | ffn_n256_pair | frame-major temporal rows (csrc/axvs.cu g_pair)


def test_pair_mode_ffn_matches(ops, O):
    """The cta_group::2 (CTA-pair) FFN kernel (bit 8, the default) against the single-CTA kernel (mask 0): same results, bit for bit."""
    p = synth.axial_layer_params(3)
    pk = ops.pack_layer({k: v.cuda() for k, v in p.items()})
    for rows in (100, 129, 5000):
        x = synth.randn(5 + rows, rows, 256)
        ref = O._ffn_tail(x, p)
        outs = []
        try:
            for pair in (0, 8):
                ops.set_pair_mode(pair)
                outs.append(ops.ln_ffn_fwd(x.cuda(), pk))
                torch.cuda.synchronize()
        finally:
            ops.set_pair_mode(PAIR_DEFAULT)
        assert all(nerr(o_, ref) < TOL for o_ in outs)
        assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("clips,T,H,W", [(1, 2, 41, 41), (3, 2, 21, 21), (2, 5, 15, 20), (1, 2, 5, 5), (9, 2, 21, 21)])
def test_pair_kernels_bit_identical_to_single_cta(ops, clips, T, H, W):
    """traj_pair_kernel, qkv_pair_kernel and ffn_n256_pair_kernel (tcgen05 cta_group::2, one M = 256 instruction stream per two tiles) run
    the arithmetic of their single-CTA counterparts in the same order: a whole axial layer must agree bit for bit, for even and odd tile
    counts (the odd tile's partner CTA works on a masked dummy tile), a single tile (falls back) and T = 5."""
    from axial_vs_b200.modules import TemporalAxialTrajectoryAttentionLayer
    layer = TemporalAxialTrajectoryAttentionLayer(256, 1024, 0.0, 0.0, "relu", 8).eval()
    layer.load_state_dict(synth.axial_layer_params(11))
    layer.cuda()
    src = synth.randn(clips * 100 + H, clips * T, H * W, 256).cuda()
    pos = synth.randn(clips * 100 + W + 1, clips, T, H, W, 256).cuda()
    outs = {}
    try:
        with torch.no_grad():
            for mask in (0, 2, 4, 8, 14, 16, PAIR_DEFAULT):
                ops.set_pair_mode(mask)
                outs[mask] = layer(src, pos)[0].clone()
                torch.cuda.synchronize()
    finally:
        ops.set_pair_mode(PAIR_DEFAULT)
    assert torch.isfinite(outs[0]).all()
    for mask, y in outs.items():
        assert torch.equal(y, outs[0]), f"pair mask {mask}"


# --------------------------------------------------------------------------------------------- clip-level decoder attention (A11)
def _bn_module_state(gz, name):
    return {k[len("bn." + name) + 1:]: torch.as_tensor(gz[k]) for k in gz.files if k.startswith("bn." + name + ".")}


@pytest.mark.parametrize("tag", ["a", "b"])
def test_decoder_attention_golden(golden, tag):
    from axial_vs_b200.decoder_attn import AttentionOperation, kmeans_cluster_update
    gz = golden(f"decoder_attn_{tag}")
    t = lambda k: torch.as_tensor(gz[k]).cuda()
    m = AttentionOperation(channels_v=256, num_heads=8).eval()
    sd = {f"{nm}.{k}": v for nm in ("_batch_norm_similarity", "_batch_norm_retrieved_value") for k, v in _bn_module_state(gz, nm).items()}
    m.load_state_dict(sd, strict=True)
    out = m.cuda()(t("q"), t("k"), t("v"))
    assert nerr(out, torch.as_tensor(gz["attn_out"])) < 1e-5          # fp32 kernel against the fp32 reference
    upd, idx = kmeans_cluster_update(t("mask_logits"), t("pixel_value"), return_assignment=True)
    assert torch.equal(idx.cpu().long(), torch.as_tensor(gz["mask_logits"]).argmax(1))       # index work: bit-exact
    assert nerr(upd, torch.as_tensor(gz["kmeans_update"])) < 1e-5


@pytest.mark.parametrize("N,L,M,advanced", [(3, 128, 3362, False), (2, 128, 13122, True), (5, 100, 882, False), (1, 7, 63, True),
                                            (32, 128, 3362, False)])
def test_kmeans_update_oracle(ops, O, N, L, M, advanced):
    g = torch.Generator().manual_seed(N * 1000 + L + M)
    logits = torch.randn(N, L, M, generator=g)
    logits[:, :, ::5] = logits[:, :, ::5].round()           # exact ties between cluster centres: first maximum must win
    logits[0, :, 1] = 0.25                                   # a pixel whose logits are all equal -> cluster 0
    pv = torch.randn(N, 256, M, generator=g)
    ref, ridx = O.kmeans_update(logits.double(), pv.double(), advanced)
    out, idx = ops.kmeans_update(logits.cuda(), pv.cuda(), advanced=advanced, return_assignment=True)
    assert torch.equal(idx.cpu().long(), ridx)
    assert nerr(out, ref.float()) < 2e-6
    # determinism and a size-independent property: without the division the update sums to the per-channel pixel total
    out2 = ops.kmeans_update(logits.cuda(), pv.cuda(), advanced=advanced)
    assert torch.equal(out, out2)
    if not advanced:
        assert torch.allclose(out.sum(-1).cpu().double(), pv.double().sum(-1), atol=1e-3, rtol=1e-4)


@pytest.mark.parametrize("N,L", [(1, 128), (4, 100), (2, 1), (3, 333)])
def test_query_self_attention_oracle(ops, O, N, L):
    g = torch.Generator().manual_seed(N + L)
    q, k = torch.randn(N, 8, 16, L, generator=g), torch.randn(N, 8, 16, L, generator=g)
    v = torch.randn(N, 8, 32, L, generator=g)
    bn = {}
    for nm, n in (("_batch_norm_similarity", 8), ("_batch_norm_retrieved_value", 256)):
        bn[nm + ".weight"] = 1.0 + 0.2 * torch.randn(n, generator=g)
        bn[nm + ".bias"] = 0.1 * torch.randn(n, generator=g)
        bn[nm + ".running_mean"] = 0.1 * torch.randn(n, generator=g)
        bn[nm + ".running_var"] = 1.0 + 0.3 * torch.rand(n, generator=g)
    ref = O.query_self_attention(q.double(), k.double(), v.double(), bn).float()
    fold = lambda nm: torch.stack([bn[nm + ".weight"] / torch.sqrt(bn[nm + ".running_var"] + 1e-3),
                                   bn[nm + ".bias"] - bn[nm + ".running_mean"] * bn[nm + ".weight"] / torch.sqrt(bn[nm + ".running_var"] + 1e-3)], 1)
    out = ops.query_self_attn(q.cuda(), k.cuda(), v.cuda(), fold("_batch_norm_similarity").contiguous().cuda(),
                              fold("_batch_norm_retrieved_value").contiguous().cuda())
    assert nerr(out, ref) < 1e-5


# --------------------------------------------------------------------------------------------- within-clip projections (f1)
@pytest.mark.parametrize("tag", ["a", "b"])
def test_projections_golden(golden, tag):
    from axial_vs_b200.projections import InputProjection, OutputProjection
    gz = golden(f"proj_{tag}")
    n, c, H, W, seed = (int(gz[k]) for k in "n c H W seed".split())
    pin, pout = synth.proj_params(seed, c)
    mi, mo = InputProjection(c).eval(), OutputProjection(2 * c).eval()
    mi.load_state_dict(pin, strict=True)
    mo.load_state_dict(pout, strict=True)
    x = synth.randn(seed + 100, n, c, H, W).cuda()
    with torch.no_grad():
        tok = mi.cuda()(x)
        y = mo.cuda()(torch.as_tensor(gz["tokens"]).cuda(), H, W)
    assert nerr(tok, torch.as_tensor(gz["tokens"])) < TOL
    assert nerr(y, torch.as_tensor(gz["y"])) < TOL


@pytest.mark.parametrize("n,c,H,W", [(4, 2048, 21, 21), (2, 1024, 41, 41), (2, 512, 81, 81), (3, 64, 5, 3),
                                     (2, 192, 33, 29), (2, 768, 21, 21), (3, 320, 11, 13)])
def test_projections_oracle_config_sizes(O, n, c, H, W):
    """R50 pyramid of BASELINE configs[1]: res5 2048 x 21^2, res4 1024 x 41^2, res3 512 x 81^2, a ragged tiny case, and channel counts
    that are not multiples of 256 on the output side (ConvNeXt-L / V2-L: res3 has 384 channels = 2 x 192; 1536 = 2 x 768)."""
    from axial_vs_b200.projections import InputProjection, OutputProjection
    seed = 700 + n + c + H
    pin, pout = synth.proj_params(seed, c)
    if (2 * c) % 32:
        pout = None
    x = synth.randn(seed + 1, n, c, H, W)
    ref_tok = O.input_proj(x, pin)
    mi = InputProjection(c).eval()
    mi.load_state_dict(pin, strict=True)
    with torch.no_grad():
        tok = mi.cuda()(x.cuda())
    assert nerr(tok, ref_tok) < TOL
    assert torch.equal(tok, mi(x.cuda()))                       # deterministic (fixed-order GroupNorm reductions)
    if pout is not None:
        mo = OutputProjection(2 * c).eval()
        mo.load_state_dict(pout, strict=True)
        with torch.no_grad():
            y = mo.cuda()(ref_tok.contiguous().cuda(), H, W)
        assert nerr(y, O.output_proj(ref_tok, pout, H, W)) < TOL


def test_projections_level_slices():
    """The projections read / write one level's slice of the multi-level token tensor [images, Len, 256] in place (what forward_features
    does instead of torch.cat / .contiguous()): bit-identical to the dense call, neighbouring levels untouched."""
    from axial_vs_b200.projections import InputProjection, OutputProjection
    n, c, shapes = 3, 128, [(5, 7), (9, 11), (6, 4)]
    pin, pout = synth.proj_params(31, c)
    mi, mo = InputProjection(c).eval().cuda(), OutputProjection(2 * c).eval().cuda()
    mi.load_state_dict(pin, strict=True)
    mo.load_state_dict(pout, strict=True)
    Len = sum(h * w for h, w in shapes)
    mem = torch.full((n, Len, 256), 7.0, device="cuda")
    start = shapes[0][0] * shapes[0][1]
    H, W = shapes[1]
    x = synth.randn(32, n, c, H, W).cuda()
    with torch.no_grad():
        dense = mi(x)
        mi(x, out=mem[:, start:start + H * W])
        assert torch.equal(mem[:, start:start + H * W], dense)
        assert bool((mem[:, :start] == 7.0).all()) and bool((mem[:, start + H * W:] == 7.0).all())
        assert torch.equal(mo(mem[:, start:start + H * W], H, W), mo(dense, H, W))


# --------------------------------------------------------------------------------------------- MSDeformAttn spatial layer (f2)
def _msda_layer(p):
    from axial_vs_b200.msda import MSDeformAttnTransformerEncoderLayer
    m = MSDeformAttnTransformerEncoderLayer(256, 1024, 0.0, "relu", 3, 8, 4).eval()
    m.load_state_dict(p, strict=True)
    return m.cuda()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_msda_layer_golden(golden, tag):
    from axial_vs_b200 import msda
    gz = golden(f"msda_layer_{tag}")
    n, seed = int(gz["n"]), int(gz["seed"])
    shapes = [tuple(int(v) for v in r) for r in gz["shapes"]]
    p = synth.msda_layer_params(seed)
    Len = sum(h * w for h, w in shapes)
    src, pos = synth.randn(seed + 100, n, Len, 256).cuda(), synth.randn(seed + 200, n, Len, 256).cuda()
    ref = msda.reference_points(shapes, n, "cuda")
    assert torch.allclose(ref.cpu(), torch.as_tensor(gz["ref_points"]), atol=1e-6)
    with torch.no_grad():
        out = _msda_layer(p)(src, pos, ref, torch.tensor(shapes), None, None)
    assert nerr(out, torch.as_tensor(gz["out"])) < TOL


@pytest.mark.parametrize("n,shapes", [(2, [(21, 21), (41, 41), (81, 81)]), (3, [(5, 7), (10, 13), (20, 27)])])
def test_msda_layer_oracle_config_sizes(O, n, shapes):
    """The R50 641x641 pyramid of BASELINE configs[1] (441 + 1681 + 6561 tokens per frame) and a ragged small pyramid."""
    from axial_vs_b200 import msda
    seed = 800 + n + shapes[0][0]
    p = synth.msda_layer_params(seed)
    Len = sum(h * w for h, w in shapes)
    src, pos = synth.randn(seed + 1, n, Len, 256), synth.randn(seed + 2, n, Len, 256)
    ref = O.msda_reference_points(shapes, n)
    want = O.msda_encoder_layer(src, pos, ref, shapes, p)
    with torch.no_grad():
        out = _msda_layer(p)(src.cuda(), pos.cuda(), msda.reference_points(shapes, n, "cuda"), shapes)
    e = nerr(out, want)
    cos = torch.nn.functional.cosine_similarity(out.cpu().flatten(), want.flatten(), dim=0).item()
    assert e < TOL and cos > 0.9999, (e, cos)


@pytest.mark.parametrize("n,shapes", [(2, [(21, 21), (41, 41), (81, 81)]), (3, [(5, 7), (10, 13), (20, 27)]), (1, [(4, 5), (7, 9), (9, 14)])])
def test_msda_front_kernel_matches_generic_gemms(ops, n, shapes):
    """The fused MSDeformAttn kernels (msda_front_pair_kernel: value + offsets | logits projections in one pass, head-major value rows;
    msda_tail_pair_kernel: output projection + residual + LayerNorm1 -> fp32 rows + the FFN's tile image) against the generic GEMMs /
    LayerNorm kernel they replace (pair-mode bit 4 off): same bf16 operands and accumulation order in the GEMMs; LayerNorm1 uses one-pass
    statistics instead of two-pass ones, so a few elements of the FFN's bf16 input round the other way -> agreement at the level of single
    bf16 flips (<= 2e-3 of the output range, measured 4e-4), far inside the 1e-2 parity tolerance.  Even / odd / tiny tile counts, with
    the positional table broadcast over the images and materialised per image."""
    from axial_vs_b200 import msda
    p = synth.msda_layer_params(77)
    Len = sum(h * w for h, w in shapes)
    src = synth.randn(78, n, Len, 256).cuda()
    pos1 = synth.randn(79, 1, Len, 256).cuda()
    ref = msda.reference_points(shapes, n, "cuda")
    layer = _msda_layer(p)
    for pos in (pos1, pos1.expand(n, -1, -1).contiguous(), None):
        with torch.no_grad():
            fused = layer(src, pos, ref[:1].contiguous(), shapes)
            prev = ops.set_pair_mode(ops.set_pair_mode(PAIR_DEFAULT) & ~4)
            try:
                generic = layer(src, pos, ref, shapes)
            finally:
                ops.set_pair_mode(prev)
        assert nerr(fused, generic) < 2e-3


def test_within_clip_encoder_golden(golden):
    """The whole within-clip transformer encoder (drop-in for MSDeformAttnTransformerEncoder): 2 stages x [MSDeformAttn spatial
    layer on 3 levels, TemporalEncoder on the first 2], same state-dict keys as the reference, against its CPU output."""
    from axial_vs_b200 import msda, modules, within_clip
    from _cases import wc_encoder_case as _wc_encoder_case
    gz = golden("wc_encoder")
    B, T, shapes, spatial, temporal_states, state, src, pos, pos3d = _wc_encoder_case(gz)
    enc = within_clip.WithinClipEncoder(msda.MSDeformAttnTransformerEncoderLayer(256, 1024, 0.0, "relu", 3, 8, 4), 2, 3, 2,
                                        modules.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "axial-trajectory", 1)).eval()
    enc.load_state_dict(state, strict=True)
    enc.cuda()
    Len = src.shape[1]
    ss = torch.tensor(shapes)
    lsi = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
    with torch.no_grad():
        out, h, w = enc(src.cuda(), ss, lsi, torch.ones(B * T, 3, 2, device="cuda"), pos.cuda(),
                        torch.zeros(B * T, Len, dtype=torch.bool, device="cuda"), [p.cuda() for p in pos3d])
    assert h is None and w is None
    assert nerr(out, torch.as_tensor(gz["out"])) < 1.5e-2        # four chained bf16 layers


class _Shape:
    def __init__(self, channels, stride):
        self.channels, self.stride = channels, stride


def _wc_module(chans):
    from axial_vs_b200 import within_clip
    shape = {"res2": _Shape(64, 4), "res3": _Shape(chans[2], 8), "res4": _Shape(chans[1], 16), "res5": _Shape(chans[0], 32)}
    return within_clip.WithinClipTrackingModule(
        shape, transformer_dropout=0.0, transformer_attn_drop=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
        transformer_num_stages=2, transformer_spatial_layers=2, transformer_temporal_layers=2,
        transformer_temporal_attn_type="axial-trajectory", conv_dims=256, transformer_spatial_in_features=["res3", "res4", "res5"],
        transformer_temporal_in_features=["res4", "res5"], num_clip_frames=2, cross_clip_training=False).eval()


def test_within_clip_module_golden(golden):
    """The whole within-clip tracking module as a drop-in for MSDeformAttnPixelDecoder: the reference's state dict loads with
    strict=True and forward_features reproduces its CPU output (projections, positional terms, 2 x [spatial + temporal layers])."""
    gz = golden("wc_module")
    seed, chans = int(gz["seed"]), [int(c) for c in gz["chans"]]
    sizes = [tuple(int(v) for v in r) for r in gz["sizes"]]
    m = _wc_module(chans)
    m.load_state_dict(synth.within_clip_module_params(seed, chans), strict=True)
    m.cuda()
    feats = {f"res{5 - i}": synth.randn(seed + 100 + i, 2, chans[i], *sizes[i]).cuda() for i in range(3)}
    out, h, w = m.forward_features(feats)
    assert h is None and w is None
    for name in ("res5", "res4", "res3"):
        want = torch.as_tensor(gz[name])
        assert out[name].shape == want.shape
        assert nerr(out[name], want) < 2e-2, name           # six chained bf16 stages, GroupNorm-normalised output


def test_within_clip_module_r50_shapes(O):
    """BASELINE configs[1] pyramid for one clip (T=2): res5 2048 x 21^2, res4 1024 x 41^2, res3 512 x 81^2, against the oracle."""
    chans, sizes, seed = [2048, 1024, 512], [(21, 21), (41, 41), (81, 81)], 123
    p = synth.within_clip_module_params(seed, chans)
    m = _wc_module(chans)
    m.load_state_dict(p, strict=True)
    m.cuda()
    feats = [synth.randn(seed + 1 + i, 2, chans[i], *sizes[i]) for i in range(3)]
    want = O.within_clip_module(feats, p, 1, 2)
    out, _, _ = m.forward_features({f"res{5 - i}": feats[i].cuda() for i in range(3)})
    for i, name in enumerate(("res5", "res4", "res3")):
        e = nerr(out[name], want[i])
        cos = torch.nn.functional.cosine_similarity(out[name].cpu().flatten(), want[i].flatten(), dim=0).item()
        assert e < 3e-2 and cos > 0.999, (name, e, cos)


def test_within_clip_module_vipseg_config_shapes(O):
    """The pyramid of the shipped VIPSeg configs (IMAGE_SIZE 769 x 1345, Vk/configs/VIPSeg/panoptic_segmentation/maxtron_wc_*.yaml) for one
    clip (T = 2): res5 2048 x 25 x 43, res4 1024 x 49 x 85, res3 512 x 97 x 169 -- rectangular maps, frames of 85 tokens in the W pass --
    against the oracle."""
    chans, sizes, seed = [2048, 1024, 512], [(25, 43), (49, 85), (97, 169)], 769
    p = synth.within_clip_module_params(seed, chans)
    m = _wc_module(chans)
    m.load_state_dict(p, strict=True)
    m.cuda()
    feats = [synth.randn(seed + 1 + i, 2, chans[i], *sizes[i]) for i in range(3)]
    want = O.within_clip_module(feats, p, 1, 2)
    out, _, _ = m.forward_features({f"res{5 - i}": feats[i].cuda() for i in range(3)})
    for i, name in enumerate(("res5", "res4", "res3")):
        e = nerr(out[name], want[i])
        cos = torch.nn.functional.cosine_similarity(out[name].cpu().flatten(), want[i].flatten(), dim=0).item()
        assert e < 3e-2 and cos > 0.999, (name, e, cos)


def test_within_clip_module_convnext_channels_multi_clip_and_graph(O):
    """ConvNeXt-L channel counts (res5 1536, res4 768, res3 384: the output side is not a multiple of 256) on a small ragged pyramid with
    TWO clips (cross-clip training layout: B = 2, T = 2): the level slices of the multi-level token tensor are written / read in place
    across four frames; against the oracle, and a CUDA-graph replay of forward_features reproduces the eager result bit for bit."""
    chans, sizes, seed = [1536, 768, 384], [(6, 7), (12, 13), (23, 26)], 321
    p = synth.within_clip_module_params(seed, chans)
    m = _wc_module(chans)
    m.cross_clip_training = True
    m.load_state_dict(p, strict=True)
    m.cuda()
    feats = [synth.randn(seed + 1 + i, 4, chans[i], *sizes[i]) for i in range(3)]
    want = O.within_clip_module(feats, p, 2, 2)
    dev = {f"res{5 - i}": feats[i].cuda() for i in range(3)}
    out, _, _ = m.forward_features(dev)
    for i, name in enumerate(("res5", "res4", "res3")):
        e = nerr(out[name], want[i])
        cos = torch.nn.functional.cosine_similarity(out[name].cpu().flatten(), want[i].flatten(), dim=0).item()
        assert e < 3e-2 and cos > 0.999, (name, e, cos)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out_g, _, _ = m.forward_features(dev)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    for name in ("res5", "res4", "res3"):
        assert torch.equal(out_g[name], out[name]), name


# --------------------------------------------------------------------------------------------- post-path tail (row f4)
def _pano(C, thr=0.3):
    from types import SimpleNamespace
    from axial_vs_b200.postprocess import PanopticPostProcessor
    thing, stuff, div = synth.panoptic_metadata(C)
    md = SimpleNamespace(thing_dataset_id_to_contiguous_id=thing, stuff_dataset_id_to_contiguous_id=stuff, label_divisor=div)
    return PanopticPostProcessor(md, pixel_confidence_threshold=thr)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_panoptic_golden(golden, tag):
    """Bit-exact panoptic ids against the unmodified reference method (fixtures made by oracle/make_golden.py)."""
    gz = golden(f"panoptic_{tag}")
    N, C, T, H, W, seed = (int(gz[k]) for k in "N C T H W seed".split())
    mc, mp, me = synth.panoptic_case(seed, N, C, T, H, W)
    seg, dic = _pano(C, float(gz["thr"])).panoptic_mask_inference(mc.cuda(), mp.cuda(), me.cuda())
    assert seg.dtype == torch.int32 and tuple(seg.shape) == (T, H, W)
    assert np.array_equal(seg.cpu().numpy(), gz["seg"])
    cats = sorted(dic.keys())
    assert cats == list(gz["cats"]) and [len(dic[c]) for c in cats] == list(gz["counts"])
    if cats:
        got = torch.cat([torch.stack(dic[c]) for c in cats]).cpu()
        assert (got - torch.from_numpy(gz["embs"])).abs().max().item() < 1e-6


@pytest.mark.parametrize("N,C,T,H,W,thr,cell", [(128, 124, 2, 161, 161, 0.3, 13), (128, 124, 2, 641, 641, 0.3, 53), (100, 19, 5, 97, 131, 0.4, 8),
                                                 (255, 50, 1, 64, 48, 0.2, 5), (1, 3, 1, 5, 7, 0.3, 4),
                                                 (120, 10, 2, 200, 160, 0.2, 1)])      # pure noise: thousands of distinct candidate sets (more than the greedy kernel keeps in registers)
def test_panoptic_oracle_sizes(N, C, T, H, W, thr, cell):
    """Config-sized inputs (128 slots, VIPSeg's 124 classes, up to the full 2 x 641 x 641 frame pair) against the numpy oracle.  Every
    decision is a comparison of fp32 scores, so the result is exact unless a score sits within rounding distance of its threshold:
    the test measures those margins on the oracle side and asserts exact equality when they are clear."""
    from oracle import panoptic_oracle as PO
    seed = 3000 + N + C + T + H + W
    mc, mp, me = synth.panoptic_case(seed, N, C, T, H, W, cell=cell)
    pp = _pano(C, thr)
    seg, segments = pp.panoptic_segments(mc.cuda(), mp.cuda())
    seg2, segments2 = pp.panoptic_segments(mc.cuda(), mp.cuda())
    assert torch.equal(seg, seg2) and torch.equal(segments, segments2), "not reproducible"
    meta = PO.Metadata(*synth.panoptic_metadata(C))
    ref, _, ref_segments = PO.panoptic_mask_inference(mc.numpy(), mp.numpy(), None, meta, pixel_thr=thr)
    px_m, cls_m, gap, top2 = PO.margins(mc.numpy(), mp.numpy(), meta, thr, 0.1, 0.3)
    got = seg.cpu().numpy()
    n = int(segments[0])
    got_segments = [tuple(r) for r in segments[1:1 + 4 * n].view(n, 4).tolist()]
    if cls_m > 1e-5 and gap > 1e-5 and top2 > 1e-5:
        mism = int((got != ref).sum())
        if px_m > 2e-6:
            assert mism == 0 and got_segments == ref_segments
        else:   # a handful of pixels within rounding distance of the pixel threshold may flip; the segment table must still agree
            assert mism <= max(2, got.size // 100000), mism
            assert [s[0] for s in got_segments] == [s[0] for s in ref_segments]
    else:
        pytest.skip(f"degenerate synthetic case (margins {cls_m:.2e} {gap:.2e} {top2:.2e})")


def test_panoptic_rejects_bad_arguments():
    mc, mp, me = synth.panoptic_case(1, 8, 6, 1, 8, 8)
    with pytest.raises(RuntimeError):
        _pano(6).panoptic_mask_inference(mc, mp, me)                                   # CPU tensors
    with pytest.raises(RuntimeError):
        _pano(6, thr=0.1).panoptic_mask_inference(mc.cuda(), mp.cuda(), me.cuda())      # more than four candidates per pixel possible
    with pytest.raises(KeyError):
        _pano(3).panoptic_mask_inference(mc.cuda(), mp.cuda(), me.cuda())               # more classes than the metadata names
    big = torch.zeros(256, 7, device="cuda")
    with pytest.raises(RuntimeError):
        _pano(6).panoptic_mask_inference(big, torch.zeros(256, 1, 4, 4, device="cuda"), torch.zeros(256, 4, device="cuda"))


# --------------------------------------------------------------------------------------------- kMaX axial attention (row f3)
def _kmax_1d(C, L, p):
    from axial_vs_b200.kmax_axial import AxialAttention
    m = AxialAttention(C, query_shape=L, total_key_depth=512, total_value_depth=1024, num_heads=8).eval()
    m.load_state_dict(p, strict=True)
    return m.cuda()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_kmax_axial_golden(golden, tag):
    gz = golden(f"kmax_axial_{tag}")
    N, C, L, seed = (int(gz[k]) for k in "N C L seed".split())
    p = synth.kmax_axial_params(seed, C)
    with torch.no_grad():
        y = _kmax_1d(C, L, p)(synth.randn(seed + 100, N, C, L).cuda())
    assert tuple(y.shape) == (N, 1024, L)
    assert nerr(y, torch.from_numpy(gz["y"])) < TOL


def test_kmax_axial_2d_golden(golden):
    from axial_vs_b200.kmax_axial import AxialAttention2D
    gz = golden("kmax_axial_2d")
    N, C, H, W, seed = (int(gz[k]) for k in "N C H W seed".split())
    m = AxialAttention2D(C, query_shape=[H, W], filters=512, key_expansion=1, value_expansion=2, num_heads=8).eval()
    m._height_axis.load_state_dict(synth.kmax_axial_params(seed, C), strict=True)
    m._width_axis.load_state_dict(synth.kmax_axial_params(seed + 1, 1024), strict=True)
    m.cuda()
    with torch.no_grad():
        y = m(synth.randn(seed + 100, N, C, H, W).cuda())
    assert tuple(y.shape) == (N, 1024, H, W)
    assert nerr(y, torch.from_numpy(gz["y"])) < TOL


@pytest.mark.parametrize("N,C,H,W", [(2, 512, 21, 21), (1, 512, 41, 41), (3, 64, 64, 5), (2, 128, 1, 33), (1, 256, 49, 85), (1, 64, 2, 110)])
def test_kmax_axial_2d_oracle_config_sizes(N, C, H, W):
    """kMaX R50 at 641x641: the axial blocks run at stride 32 (21x21) and stride 16 (41x41) on 512 channels; the stride-16 map of the shipped
    769 x 1345 VIPSeg configs (49 x 85); plus a 110-long axis (the limit at the default depths is 112) and a degenerate one."""
    from axial_vs_b200.kmax_axial import AxialAttention2D
    from oracle import kmax_oracle as KO
    seed = 4000 + N + C + H + W
    ph, pw = synth.kmax_axial_params(seed, C), synth.kmax_axial_params(seed + 1, 1024)
    m = AxialAttention2D(C, query_shape=[H, W]).eval()
    m._height_axis.load_state_dict(ph, strict=True)
    m._width_axis.load_state_dict(pw, strict=True)
    m.cuda()
    x = synth.randn(seed + 2, N, C, H, W)
    ref = KO.axial_attention_2d(x, ph, pw)
    with torch.no_grad():
        y = m(x.cuda())
    e = nerr(y, ref)
    cos = torch.nn.functional.cosine_similarity(y.cpu().flatten(), ref.flatten(), dim=0).item()
    assert e < TOL and cos > 0.9999, (e, cos)


def test_kmax_axial_rejects_bad_arguments():
    p = synth.kmax_axial_params(1, 64)
    m = _kmax_1d(64, 9, p)
    with torch.no_grad():
        with pytest.raises(RuntimeError):
            m(torch.zeros(1, 64, 9))                      # CPU tensor
        with pytest.raises(RuntimeError):
            m(torch.zeros(1, 64, 129, device="cuda"))     # axis longer than the kernel supports (128)
        with pytest.raises(RuntimeError):
            m(torch.zeros(1, 32, 9, device="cuda"))       # wrong channel count
    m.train()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 64, 9, device="cuda"))


def test_kmax_axial_tensor_core_and_simt_paths_agree():
    """The split-bf16 mma.sync attention core against the fp32 SIMT kernel (same GEMM in front): both fp32-grade."""
    from axial_vs_b200 import _lib
    from axial_vs_b200.kmax_axial import AxialAttention2D
    N, C, H, W = 2, 128, 41, 23
    m = AxialAttention2D(C, query_shape=[H, W]).eval()
    m._height_axis.load_state_dict(synth.kmax_axial_params(71, C), strict=True)
    m._width_axis.load_state_dict(synth.kmax_axial_params(72, 1024), strict=True)
    m.cuda()
    x = synth.randn(73, N, C, H, W).cuda()
    lib = _lib.load()
    try:
        with torch.no_grad():
            lib.axvs_set_kmax_tensor_cores(0)
            a = m(x)
            lib.axvs_set_kmax_tensor_cores(1)
            b = m(x)
        torch.cuda.synchronize()
    finally:
        lib.axvs_set_kmax_tensor_cores(1)
    assert nerr(b, a) < 1e-4


# --------------------------------------------------------------------------------------------- end-to-end labels (row N1)
def test_label_agreement_end_to_end(O):
    """GPU logits -> GPU panoptic merge against reference (oracle) logits -> the reference's merge (numpy restatement, pinned on the
    unmodified method), class thresholds overridden to 0 as SURVEY.md 8c(ii) prescribes.  Random-init logits are nearly flat (every
    softmax confidence ~ 1/128 < the 0.3 pixel threshold -> everything void), so both sides' mask logits are multiplied by the same
    temperature, which makes the confidences decisive the way a trained head does without touching the comparison."""
    from axial_vs_b200 import cross_clip
    from oracle import panoptic_oracle as PO
    Q, T, V, H, W, L, K, seed = 128, 4, 2, 48, 48, 4, 124, 4242
    p = synth.cross_clip_params(seed, L, K)
    m = cross_clip.CrossClipTrackingModule(num_layers=L, num_classes=K, attn_drop=0.0, aspp_drop=0.0, kernel_sizes=[3, 3, 3],
                                           atrous_rates=[1, 2, 3], norm_fn="ln", num_clip_frames=V).eval()
    m.load_state_dict(p, strict=True)
    m.cuda()
    cq = synth.randn(seed + 1, 1, Q, T, 256)
    pf = synth.randn(seed + 2, 1, 128, T * V, H, W)
    ref = O.cross_clip_module(cq, pf, p, L, V)
    with torch.no_grad():
        o = m(cq.cuda(), pf.cuda())
    got_masks, want_masks = o["pred_masks"][0].float(), ref["pred_masks"][0].float()          # [Q, T*V, H, W]
    agree_q = (got_masks.cpu().argmax(0) == want_masks.argmax(0)).float().mean().item()
    temp = 60.0
    pp = _pano(K, 0.3)
    pp.class_threshold_thing = pp.class_threshold_stuff = 0.0
    seg, segments = pp.panoptic_segments(o["pred_logits"][0], got_masks * temp)
    meta = PO.Metadata(*synth.panoptic_metadata(K))
    ref_seg, _, ref_segments = PO.panoptic_mask_inference(ref["pred_logits"][0].numpy(), (want_masks * temp).numpy(), None, meta, pixel_thr=0.3,
                                                          thing_thr=0.0, stuff_thr=0.0)
    seg = seg.cpu().numpy()
    n = int(segments[0])
    got_slots = {(r[0], r[1]) for r in segments[1:1 + 4 * n].view(n, 4).tolist()}              # (slot, label) of every accepted segment
    ref_slots = {(r[0], r[1]) for r in ref_segments}
    jacc = len(got_slots & ref_slots) / max(1, len(got_slots | ref_slots))
    assigned = float((ref_seg >= 0).mean())
    agree_p = float((seg == ref_seg).mean())
    div = meta.label_divisor
    agree_c = float(((seg // div) == (ref_seg // div)).mean())                                  # category of the pixel (instance ids ignored)
    print(f"[labels end to end] query argmax agreement {agree_q:.5f}; panoptic: {assigned:.1%} of the pixels assigned, {len(ref_slots)} segments, "
          f"segment-set Jaccard {jacc:.3f}, pixel agreement ids {agree_p:.5f} / categories {agree_c:.5f}")
    assert assigned > 0.3, "the synthetic case must exercise the merge"
    # The greedy merge is a cascade of discrete decisions (slot order by score, overlap test at 0.8): on these near-tied random-init logits
    # one flipped accept / reject re-labels every pixel of a slot, so id agreement sits well below the per-pixel query agreement.
    assert agree_q >= 0.998 and jacc >= 0.85 and agree_c >= 0.995      # measured 0.9985, 0.93, 0.9983 (instance ids renumber after one flipped slot: 0.966)
    # the merge itself, on IDENTICAL (reference) logits, is exact: the panoptic kernels decide by comparisons of fp32 scores
    seg_same, _ = pp.panoptic_segments(ref["pred_logits"][0].cuda(), (want_masks * temp).cuda())
    assert float((seg_same.cpu().numpy() == ref_seg).mean()) >= 0.999


def test_label_agreement_split_precision_cross_clip(O):
    """north_star asks for >= 99.9 % per-pixel agreement of the final argmax labels.  With the cross-clip trajectory attention at fp32-grade
    accuracy (`set_precision("split")`: split-precision GEMMs + fp32 attention kernels) the GPU labels meet it on the same near-tied
    random-init logits on which the bf16 attention reaches 99.85 % (test above); the attention output itself is then fp32-grade too."""
    from axial_vs_b200 import cross_clip
    Q, T, V, H, W, L, K, seed = 128, 4, 2, 48, 48, 4, 124, 4242
    p = synth.cross_clip_params(seed, L, K)
    m = cross_clip.CrossClipTrackingModule(num_layers=L, num_classes=K, attn_drop=0.0, aspp_drop=0.0, kernel_sizes=[3, 3, 3],
                                           atrous_rates=[1, 2, 3], norm_fn="ln", num_clip_frames=V).eval()
    m.load_state_dict(p, strict=True)
    m.cuda().set_precision("split")
    cq = synth.randn(seed + 1, 1, Q, T, 256)
    pf = synth.randn(seed + 2, 1, 128, T * V, H, W)
    ref = O.cross_clip_module(cq, pf, p, L, V)
    with torch.no_grad():
        o = m(cq.cuda(), pf.cuda())
    got, want = o["pred_masks"][0].float().cpu(), ref["pred_masks"][0].float()
    agree = (got.argmax(0) == want.argmax(0)).float().mean().item()
    print(f"[labels, split-precision cross-clip attention] query argmax agreement {agree:.5f}, logit error {nerr(got, want):.2e}")
    assert nerr(got, want) < 1e-3
    assert agree >= 0.999, agree
    # one layer on its own: fp32-grade against the oracle
    layer = m.transformer_trajectory_self_attention_layers[0]
    x = synth.randn(seed + 3, 2, T * Q, 256)
    lp = {k[len("transformer_trajectory_self_attention_layers.0."):]: v for k, v in p.items() if k.startswith("transformer_trajectory_self_attention_layers.0.")}
    with torch.no_grad():
        y = layer(x.cuda(), seq_len=Q, num_frames=T)
    assert nerr(y, O.cc_attention_layer(x, lp, Q, T)) < 2e-4


# --------------------------------------------------------------------------------------------- clip-to-clip matching (row f4)
def _lsap_cases():
    rng = np.random.default_rng(7)
    cases = []
    for k in range(18):
        for n in (4, 31, 64, 100, 128):
            kind = k % 6
            if kind == 0:
                c = rng.random((n, n))
            elif kind == 1:
                c = rng.integers(0, 3, (n, n))                       # heavy ties
            elif kind == 2:
                c = rng.integers(0, 20, (n, n))
            elif kind == 3:
                c = np.full((n, n), 0.5)                             # constant: scipy returns the identity
            elif kind == 4:
                e = rng.standard_normal((n, 32))
                e[n // 2] = e[0]                                     # duplicated embedding: exact cosine ties
                en = e / np.linalg.norm(e, axis=1, keepdims=True)
                c = 1 - en @ en[rng.permutation(n)].T
            else:
                c = rng.random((n, n)) - 0.5
            cases.append(np.asarray(c, dtype=np.float32))
    cases.append(rng.random((256, 256)).astype(np.float32))
    cases.append(rng.integers(0, 4, (256, 256)).astype(np.float32))
    return cases


def test_lsap_matches_scipy_including_ties():
    """axvs_lsap against scipy.optimize.linear_sum_assignment (the reference's call, maxtron_wc_model.py:398) on 92 seeded matrices:
    random, integer-valued (heavily tied), constant, tied cosine costs, n up to 256.  Same algorithm, arithmetic and tie rules -> the
    permutations are identical, not merely equally good."""
    from scipy.optimize import linear_sum_assignment
    from axial_vs_b200 import matching
    cases = _lsap_cases()
    by_n = {}
    for c in cases:
        by_n.setdefault(c.shape[0], []).append(c)
    n_checked = 0
    for n, cs in by_n.items():
        got = matching.linear_sum_assignment(torch.from_numpy(np.stack(cs)).cuda()).cpu().numpy()      # one launch per size, one CTA per matrix
        for c, g_ in zip(cs, got):
            want = linear_sum_assignment(c)[1]
            assert np.array_equal(g_, want), f"n={n}"
            n_checked += 1
    assert n_checked >= 90


@pytest.mark.parametrize("videos,clips,n,e", [(1, 6, 128, 128), (3, 9, 100, 256), (2, 2, 17, 40)])
def test_match_chain_against_reference_loop(videos, clips, n, e):
    """The whole chain (normalise, cosine cost, assignment, permute, next clip) in ONE launch per call against the reference's loop
    restated with torch + scipy (oracle/matching_oracle.py), without a host synchronisation inside the call."""
    from axial_vs_b200 import matching
    from oracle import matching_oracle as MO
    g = torch.Generator().manual_seed(videos * 100 + clips)
    emb = torch.randn(videos, clips, n, e, generator=g)
    emb[:, 1:] = 0.7 * emb[:, 1:] + 0.3 * emb[:, :1][:, :, torch.randperm(n, generator=g)]      # clips share structure, shuffled
    x = emb.cuda()
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        idx = matching.match_chain(x)                                   # must not synchronise with the host
    finally:
        torch.cuda.set_sync_debug_mode("default")
    idx = idx.cpu().numpy()
    for v in range(videos):
        assert np.array_equal(idx[v], MO.match_chain(emb[v]))
    # reference-signature wrapper
    one = matching.match_from_embds(x[0, 0], x[0, 1]).cpu().numpy()
    assert np.array_equal(one, MO.match_from_embds(emb[0, 0], emb[0, 1]))


# --------------------------------------------------------------------------------------------- Tube-Link rows (A7, A10)
def test_tube_link_temporal_encoder_golden(golden):
    """Tube-Link TemporalEncoder drop-in against the output of the UNMODIFIED Tube-Link classes (tests/golden/tl_temporal.npz)."""
    from axial_vs_b200 import tube_link
    gz = golden("tl_temporal")
    B, T, H, W, seed = (int(gz[k]) for k in "B T H W seed".split())
    enc = tube_link.TemporalEncoder(256, 1024, attn_drop=0.0, num_temporal_layer=1).eval()
    enc.load_state_dict(synth.encoder_params(seed, 1), strict=True)
    enc.cuda()
    src = synth.randn(seed + 1, B * T, H * W, 256)
    from oracle import traj_oracle as O_
    pos = O_.level_pos3d(B, T, H, W, synth.level_embed(seed + 2)[0])
    with torch.no_grad():
        y = enc(src=src.cuda(), pos=pos.cuda())
    assert nerr(y, torch.from_numpy(gz["y"])) < TOL


def test_tube_link_cc_layer_golden(golden):
    from axial_vs_b200 import cross_clip
    gz = golden("tl_cc_layer")
    b, Q, T, seed = (int(gz[k]) for k in "b Q T seed".split())
    p = {}
    g = torch.Generator().manual_seed(seed)
    synth.traj_attn_params(g, "self_attn.", 256, p, fused_qkv=True)
    p["norm.weight"] = 1 + 0.1 * torch.randn(256, generator=g)
    p["norm.bias"] = 0.1 * torch.randn(256, generator=g)
    m = cross_clip.TrajectoryAttentionLayer(256, 8).eval()
    m.load_state_dict(p, strict=True)
    m.cuda()
    with torch.no_grad():
        y = m(synth.randn(seed + 1, b, T * Q, 256).cuda(), seq_len=Q, num_frames=T)
    assert nerr(y, torch.from_numpy(gz["y"])) < TOL


def _tl_attention_params(seed, n_levels, n_temporal_layers=1):
    g = torch.Generator().manual_seed(seed)
    p = {k[len("self_attn."):]: v for k, v in synth.msda_layer_params(seed, n_levels).items() if k.startswith("self_attn.")}
    p.update({f"temporal_layer.{k}": v for k, v in synth.encoder_params(seed + 1, n_temporal_layers).items()})
    p["gamma"] = 0.5 + torch.rand(256, generator=g)
    return p


@pytest.mark.parametrize("batch_first", [True, False])
def test_tube_link_axial_trajectory_msda(O, batch_first):
    """MultiScaleDeformableAxialTrajectoryAttention (TL plugin :393-638): sampling -> gamma-skip temporal branch on the two
    low-resolution levels -> shared output_proj -> identity residual, against the oracle composition (whose temporal classes are
    pinned on the Tube-Link sources by tl_temporal.npz)."""
    from axial_vs_b200 import tube_link, msda
    B, T, seed = 2, 3, 6100
    shapes = [(5, 6), (9, 11), (17, 21)]
    p = _tl_attention_params(seed, len(shapes))
    m = tube_link.MultiScaleDeformableAxialTrajectoryAttention(256, 8, num_levels=3, num_temporal_levels=2, num_temporal_layers=1, num_temporal_dim=1024,
                                                               num_points=4, dropout=0.0, batch_first=batch_first).eval()
    m.load_state_dict(p, strict=True)
    m.cuda()
    n = sum(h * w for h, w in shapes)
    query = synth.randn(seed + 2, B * T, n, 256)
    qpos = synth.randn(seed + 3, B * T, n, 256)
    le = synth.level_embed(seed + 4)
    pos3d = [O.level_pos3d(B, T, h, w, le[i]) for i, (h, w) in enumerate(shapes[:2])]
    ref = O.msda_reference_points(shapes, B * T)
    want = O.tl_axial_trajectory_msda(query, qpos, pos3d, ref, shapes, p, 2)
    ss = torch.tensor(shapes)
    if batch_first:
        got = m(query.cuda(), query_pos=qpos.cuda(), query_pos3d=[x.cuda() for x in pos3d], reference_points=ref.cuda(), spatial_shapes=ss)
    else:
        got = m(query.permute(1, 0, 2).cuda(), query_pos=qpos.permute(1, 0, 2).cuda(), query_pos3d=[x.cuda() for x in pos3d],
                reference_points=ref.cuda(), spatial_shapes=ss).permute(1, 0, 2)
    assert nerr(got, want) < TOL


def test_tube_link_forward_head_clips(O):
    """forward_head_clips + pred_class of the Tube-Link cross-clip head (TL cc head :761-797): 100 queries, 256-channel mask features."""
    from axial_vs_b200 import tube_link
    t, l, q, c, K, fpc, h, w, seed = 4, 3, 100, 256, 40, 2, 24, 20, 6200
    g = torch.Generator().manual_seed(seed)
    head = tube_link.CCHeadPredictor(256, 256, K).eval()
    with torch.no_grad():
        for prm in head.parameters():
            prm.copy_(torch.randn(prm.shape, generator=g) * (0.06 if prm.dim() > 1 else 0.1))
        head.post_norm.weight.add_(1.0)
    p = {k: v.detach().clone() for k, v in head.state_dict().items()}
    head.cuda()
    dec = synth.randn(seed + 1, t, l, q, 1, c)
    mf = synth.randn(seed + 2, 1, t * fpc, c, h, w)
    want_cls, want_masks = O.tl_forward_head_clips(dec, mf, p)
    with torch.no_grad():
        cls, masks = head.forward_head_clips(dec.cuda(), mf.cuda())
    assert len(cls) == l and len(masks) == l
    for i in range(l):
        assert nerr(cls[i], want_cls[i]) < TOL
        assert nerr(masks[i], want_masks[i]) < TOL
        agree = (masks[i].cpu().argmax(2) == want_masks[i].argmax(2)).float().mean().item()
        assert agree >= 0.998, agree


# --------------------------------------------------------------------------------------------- Tube-Link mask decoder layer (row A11)
@pytest.mark.parametrize("Nq,B,L", [(100, 2, 240), (100, 1, 1500), (100, 1, 6000), (37, 3, 70), (130, 1, 200)])
def test_tube_link_decoder_layer(golden, Nq, B, L):
    """DetrTransformerDecoderLayer drop-in (masked cross-attention over T*h*w keys -> LN -> query self-attention -> LN -> FFN -> LN) against
    the oracle restatement (pinned on torch.nn.MultiheadAttention; the mmcv wrapper itself is restated, "parity unpinned") and, at the golden
    shape, against the stored output of the torch-module composition.  Key counts of the cfg4 pyramid (1500, 6000), ragged sizes (key
    count not a multiple of 16 or 64, more than 128 queries), a query row with a single visible key."""
    from axial_vs_b200 import tube_link
    from oracle import tl_decoder_oracle as TO
    seed = 7300
    p = synth.tl_decoder_layer_params(seed)
    layer = tube_link.DetrTransformerDecoderLayer(256, 8, 2048).eval()
    layer.load_state_dict(p, strict=True)
    layer.cuda()
    q, qp, k, kp, m = synth.tl_decoder_case(seed + 1, Nq, B, L)
    want = TO.decoder_layer(q, k, k, qp, kp, [m, None], p)
    with torch.no_grad():
        got = layer(query=q.cuda(), key=k.cuda(), value=k.cuda(), query_pos=qp.cuda(), key_pos=kp.cuda(), attn_masks=[m.cuda(), None])
    assert got.shape == want.shape and nerr(got, want) < TOL
    if (Nq, B, L) == (100, 2, 240):
        assert nerr(got, torch.from_numpy(golden("tl_decoder_layer")["y"])) < TOL


def test_masked_mha_core_fp32():
    """The attention core alone (axvs_masked_mha_fwd, fp32 SIMT, flash-decoding split over the keys) against a float64 softmax: the split /
    combine must be exact to fp32 rounding, including a row whose visible keys all sit in the LAST split."""
    from axial_vs_b200 import ops
    g = torch.Generator().manual_seed(11)
    Nq, B, L, H = 100, 2, 3000, 8
    q, k, v = (torch.randn(n, B, 256, generator=g) for n in (Nq, L, L))
    mask = torch.rand(B * H, Nq, L, generator=g) < 0.5
    mask[3, 5, :L - 4] = True
    mask[3, 5, L - 4:] = False
    scale = 32 ** -0.5
    qh = q.double().reshape(Nq, B, H, 32).permute(1, 2, 0, 3) * scale
    kh = k.double().reshape(L, B, H, 32).permute(1, 2, 0, 3)
    vh = v.double().reshape(L, B, H, 32).permute(1, 2, 0, 3)
    s = (qh @ kh.transpose(-1, -2)).masked_fill(mask.reshape(B, H, Nq, L), float("-inf"))
    want = (torch.softmax(s, -1) @ vh).permute(2, 0, 1, 3).reshape(Nq, B, 256).float()
    got = ops.masked_mha((q * (scale * 1.4426950408889634)).cuda(), k.cuda(), v.cuda(), mask.cuda(), heads=H, seq_first=True, out_dtype=torch.float32)
    assert float((got.cpu() - want).abs().max()) < 2e-5


# --------------------------------------------------------------------------------------------- clip-level kMaX decoder layer, whole (row A11)
def _kmax_layer_check(got_q, got_pred, want_q, want_pred, tol_q=2e-3):
    # split-precision GEMMs: fp32-grade, so the tolerances sit far below the path's 1e-2; the k-means step is an argmax over the mask
    # logits, so a pixel whose two best clusters are closer than the logit error may move: bounded through the assignment agreement
    for name in ("class_logits", "mask_logits", "mask_embeddings", "pixel_feature"):
        assert got_pred[name].shape == want_pred[name].shape, name
        assert nerr(got_pred[name], want_pred[name]) < 2e-4, (name, nerr(got_pred[name], want_pred[name]))
    agree = (got_pred["mask_logits"].cpu().argmax(1) == want_pred["mask_logits"].argmax(1)).float().mean().item()
    assert agree >= 0.999, agree
    assert got_q.shape == want_q.shape and nerr(got_q, want_q) < tol_q, nerr(got_q, want_q)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_kmax_transformer_layer_golden(golden, tag):
    """kMaXTransformerLayer drop-in (pixel side: GELU -> 1x1 ConvBN -> value / depthwise 5x5 + two 1x1 ConvBN + L2 normalise; query x pixel
    contraction + BN; k-means update; query self-attention; FFN) against the output of the unmodified reference layer."""
    from axial_vs_b200 import decoder_attn
    gz = golden(f"kmax_layer_{tag}")
    N, L, Cp, TH, W, K, seed = (int(gz[k]) for k in "N L Cp TH W K seed".split())
    layer = decoder_attn.kMaXTransformerLayer(num_classes=K, in_channel_pixel=Cp).eval()
    layer.load_state_dict(synth.kmax_layer_params(seed, Cp, K), strict=True)
    layer.cuda()
    with torch.no_grad():
        q, pred = layer(synth.randn(seed + 100, N, Cp, TH, W).cuda(), synth.randn(seed + 200, N, 256, L).cuda())
    want = {k: torch.from_numpy(gz[k]) for k in ("class_logits", "mask_logits", "mask_embeddings", "pixel_feature")}
    _kmax_layer_check(q, pred, torch.from_numpy(gz["query"]), want)


def test_kmax_transformer_layer_oracle_r50_stage():
    """The same at a BASELINE-sized stage: 3 clips of T = 2 frames at output stride 16 (2 x 41 x 41 pixels stacked along H), 1024 pixel channels,
    128 queries, VIPSeg's 124 + 1 classes, against the oracle (pinned on the reference by the goldens)."""
    from axial_vs_b200 import decoder_attn
    from oracle import kmax_layer_oracle as KO
    N, L, Cp, TH, W, K, seed = 3, 128, 1024, 82, 41, 125, 8200
    p = synth.kmax_layer_params(seed, Cp, K)
    layer = decoder_attn.kMaXTransformerLayer(num_classes=K, in_channel_pixel=Cp).eval()
    layer.load_state_dict(p, strict=True)
    layer.cuda()
    pf, qf = synth.randn(seed + 100, N, Cp, TH, W), synth.randn(seed + 200, N, 256, L)
    want_q, want_pred = KO.transformer_layer(pf, qf, p)
    with torch.no_grad():
        q, pred = layer(pf.cuda(), qf.cuda())
    _kmax_layer_check(q, pred, want_q, want_pred)
