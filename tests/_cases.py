"""Shared construction of multi-module test cases (imported by the CPU and the GPU test files)."""
import torch  # noqa: F401

from axial_vs_b200 import synth
from oracle import traj_oracle as O


def wc_encoder_case(gz):
    B, T, seed = int(gz["B"]), int(gz["T"]), int(gz["seed"])
    shapes = [tuple(int(v) for v in r) for r in gz["shapes"]]
    spatial = [synth.msda_layer_params(seed + i) for i in range(2)]
    temporal_states = [synth.encoder_params(seed + 10 + i, 1) for i in range(2)]
    state = {}
    for i in range(2):
        state.update({f"spatial_layers.{i}.{k}": v for k, v in spatial[i].items()})
        state.update({f"temporal_layers.{i}.{k}": v for k, v in temporal_states[i].items()})
    Len = sum(h * w for h, w in shapes)
    src, pos = synth.randn(seed + 100, B * T, Len, 256), synth.randn(seed + 200, B * T, Len, 256)
    le = synth.level_embed(seed + 300)
    pos3d = [O.level_pos3d(B, T, h, w, le[i]) for i, (h, w) in enumerate(shapes[:2])]
    return B, T, shapes, spatial, temporal_states, state, src, pos, pos3d
