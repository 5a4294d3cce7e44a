"""TEST INFRASTRUCTURE ONLY -- Tube-Link goldens from the UNMODIFIED Tube-Link sources (run where /root/reference is mounted).

The Tube-Link hot-path classes are exec'd as line slices of the reference files (oracle/ref_loader.tl_*: the files import mmcv at module
level, the classes themselves are pure torch + einops), fed with the seeded synthetic weights / inputs of axial_vs_b200/synth.py, and their
outputs are stored under tests/golden/:
    tl_temporal.npz   TemporalEncoder(1 axial-trajectory layer) of TL/mmdet/models/plugins/msdeformattn_pixel_decoder.py:711-791
    tl_cc_layer.npz   TrajectoryAttentionLayer of TL/models/video/tube_link_vis/mask2former_video_cc_head.py:152-247
usage: python oracle/make_golden_tl.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from axial_vs_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle import traj_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    if not ref_loader.tl_available():
        raise SystemExit("the Tube-Link reference tree is not mounted")
    torch.manual_seed(0)
    _, TemporalEncoder, _ = ref_loader.tl_temporal_classes()
    B, T, H, W, seed = 2, 3, 6, 5, 4100
    p = synth.encoder_params(seed, 1)
    enc = TemporalEncoder(256, 1024, attn_drop=0.0, num_temporal_layer=1).eval()
    enc.load_state_dict(p, strict=True)
    src = synth.randn(seed + 1, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 2)[0])
    with torch.no_grad():
        y = enc(src=src, pos=pos)
    np.savez_compressed(os.path.join(GOLD, "tl_temporal.npz"), B=B, T=T, H=H, W=W, seed=seed, wsum=synth.checksum(p), y=y.numpy())

    Layer, _ = ref_loader.tl_cc_classes()
    b, Q, Tc, seed = 1, 12, 4, 4200
    p = {}
    g = torch.Generator().manual_seed(seed)
    synth.traj_attn_params(g, "self_attn.", 256, p, fused_qkv=True)
    p["norm.weight"] = 1 + 0.1 * torch.randn(256, generator=g)
    p["norm.bias"] = 0.1 * torch.randn(256, generator=g)
    m = Layer(256, 8).eval()
    m.load_state_dict(p, strict=True)
    x = synth.randn(seed + 1, b, Tc * Q, 256)
    with torch.no_grad():
        y = m(x, seq_len=Q, num_frames=Tc)
    np.savez_compressed(os.path.join(GOLD, "tl_cc_layer.npz"), b=b, Q=Q, T=Tc, seed=seed, wsum=synth.checksum(p), y=y.numpy())
    print("wrote tl_temporal.npz, tl_cc_layer.npz")


if __name__ == "__main__":
    main()
