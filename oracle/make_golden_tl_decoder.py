"""Golden for the Tube-Link mask decoder layer (row A11).  mmcv-full 1.6.1 is not installed / vendored, so the generator is the layer
composed from the torch modules mmcv wraps (torch.nn.MultiheadAttention, LayerNorm, Linear) with mmcv 1.x's wrapper semantics
(oracle/tl_decoder_oracle.py::torch_module_composition).  Run: python oracle/make_golden_tl_decoder.py -> tests/golden/tl_decoder_layer.npz"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from axial_vs_b200 import synth
from oracle import tl_decoder_oracle as TO

seed, Nq, B, L = 7300, 100, 2, 5 * 6 * 8          # 100 queries, T*h*w = 240 keys
p = synth.tl_decoder_layer_params(seed)
q, qp, k, kp, m = synth.tl_decoder_case(seed + 1, Nq, B, L)
y = TO.torch_module_composition(q, k, k, qp, kp, [m, None], p)
out = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "tl_decoder_layer.npz")
np.savez_compressed(out, y=y.numpy(), seed=seed, Nq=Nq, B=B, L=L, generator="torch " + torch.__version__)
print("wrote", out, y.shape, float(y.abs().max()))
