"""Which stage of WithinClipTrackingModule.forward_features breaks a CUDA-graph capture?  (debug aid)"""
import sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import synth, within_clip, ops, msda
class S:
    def __init__(self, c, s): self.channels, self.stride = c, s
chans, sizes = [1536, 768, 384], [(6, 7), (12, 13), (23, 26)]
shape = {"res2": S(256, 4), "res3": S(chans[2], 8), "res4": S(chans[1], 16), "res5": S(chans[0], 32)}
m = within_clip.WithinClipTrackingModule(shape, transformer_dropout=0.0, transformer_attn_drop=0.0, transformer_nheads=8,
        transformer_dim_feedforward=1024, transformer_num_stages=2, transformer_spatial_layers=2, transformer_temporal_layers=2,
        transformer_temporal_attn_type="axial-trajectory", conv_dims=256, transformer_spatial_in_features=["res3", "res4", "res5"],
        transformer_temporal_in_features=["res4", "res5"], num_clip_frames=2, cross_clip_training=True).eval()
m.load_state_dict(synth.within_clip_module_params(1, chans), strict=True)
m.cuda()
feats = {f"res{5 - i}": torch.randn(4, chans[i], *sizes[i], device="cuda") for i in range(3)}
names = ["res5", "res4", "res3"]
Len = sum(h * w for h, w in sizes)
with torch.no_grad():
    m.forward_features(feats); torch.cuda.synchronize()
    src = torch.randn(4, Len, 256, device="cuda")
    enc = m.transformer.encoder
    pos2d = m._pos2d(sizes, src.device)
    pos3d = [m.pe_layer_3d.table(2, 2, h, w, src.device, m.transformer.level_embed_3d[i]) for i, (h, w) in enumerate(sizes[:2])]
    ref = enc._ref_cache[(tuple(sizes), str(src.device))]
    stages = {
        "input_proj": lambda: m.input_proj[2](feats["res3"]),
        "input_proj(out=slice)": lambda: m.input_proj[1](feats["res4"], out=src[:, 42:42 + 156]),
        "pos3d table": lambda: m.pe_layer_3d.table(2, 2, 6, 7, src.device, m.transformer.level_embed_3d[0]),
        "msda layer": lambda: enc.spatial_layers[0](src, pos2d, ref, sizes),
        "temporal layer (1 level)": lambda: enc.temporal_layers[0](src=src[:, :42].contiguous(), pos=pos3d[0]),
        "temporal levels concurrent": lambda: within_clip.run_temporal_levels(enc.temporal_layers[0], src.clone(), sizes, pos3d, 2, inplace=True),
        "encoder": lambda: enc(src, sizes, None, None, pos2d, None, pos3d),
        "output_proj": lambda: m.output_proj[2](src[:, 42 + 156:], 23, 26),
        "forward_features": lambda: m.forward_features(feats),
    }
    for name, fn in stages.items():
        fn(); torch.cuda.synchronize()
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            g.replay(); torch.cuda.synchronize()
            print(f"{name:32s} capture OK", flush=True)
        except Exception as e:
            print(f"{name:32s} capture FAILED: {type(e).__name__}: {str(e).splitlines()[0]}", flush=True)
            torch.cuda.synchronize()
