"""Goldens of the clip-level kMaX decoder layer (row A11) from the UNMODIFIED reference kMaXTransformerLayer (eval mode, CPU fp32).
Run where /root/reference is mounted: python oracle/make_golden_kmax_layer.py -> tests/golden/kmax_layer_{a,b}.npz"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from axial_vs_b200 import synth
from oracle import ref_loader

CASES = {"a": (2, 128, 64, 10, 7, 20, 8101), "b": (1, 37, 128, 6, 5, 11, 8102)}     # N, L, C_pixel, TH, W, classes, seed


def reference_layer(Cp, K, seed):
    ref_loader.cross_clip()
    DEC = sys.modules["maxtron_deeplab.modeling.transformer_decoder.maxtron_transformer_decoder"]
    layer = DEC.kMaXTransformerLayer(num_classes=K, in_channel_pixel=Cp, in_channel_query=256, base_filters=128, num_heads=8, bottleneck_expansion=2,
                                     key_expansion=1, value_expansion=2).eval()
    missing = layer.load_state_dict(synth.kmax_layer_params(seed, Cp, K), strict=False)       # only the num_batches_tracked counters are absent
    assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing
    return layer


if __name__ == "__main__":
    for tag, (N, L, Cp, TH, W, K, seed) in CASES.items():
        layer = reference_layer(Cp, K, seed)
        pf, qf = synth.randn(seed + 100, N, Cp, TH, W), synth.randn(seed + 200, N, 256, L)
        with torch.no_grad():
            q, pred = layer(pf, qf)
        out = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"kmax_layer_{tag}.npz")
        np.savez_compressed(out, N=N, L=L, Cp=Cp, TH=TH, W=W, K=K, seed=seed, query=q.numpy(), class_logits=pred["class_logits"].numpy(),
                            mask_logits=pred["mask_logits"].numpy(), mask_embeddings=pred["mask_embeddings"].numpy(),
                            pixel_feature=pred["pixel_feature"].numpy())
        print("wrote", out, os.path.getsize(out))
