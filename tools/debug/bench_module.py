"""Time the whole within-clip tracking module (WithinClipTrackingModule.forward_features) at the R50 641x641 pyramid."""
import sys, time, torch
sys.path.insert(0, ".")
from axial_vs_b200 import synth, within_clip, ops
clips = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = 2
class S:
    def __init__(self, c, s): self.channels, self.stride = c, s
chans, sizes = [2048, 1024, 512], [(21, 21), (41, 41), (81, 81)]
shape = {"res2": S(256, 4), "res3": S(512, 8), "res4": S(1024, 16), "res5": S(2048, 32)}
m = within_clip.WithinClipTrackingModule(shape, transformer_dropout=0.0, transformer_attn_drop=0.0, transformer_nheads=8,
        transformer_dim_feedforward=1024, transformer_num_stages=2, transformer_spatial_layers=2, transformer_temporal_layers=4,
        transformer_temporal_attn_type="axial-trajectory", conv_dims=256, transformer_spatial_in_features=["res3", "res4", "res5"],
        transformer_temporal_in_features=["res4", "res5"], num_clip_frames=T, cross_clip_training=True).eval()
m.load_state_dict(synth.within_clip_module_params(1, chans, 2, 2), strict=True)
m.cuda()
feats = {f"res{5 - i}": torch.randn(clips * T, chans[i], *sizes[i], device="cuda") for i in range(3)}
for _ in range(2): m.forward_features(feats)
torch.cuda.synchronize()
ops.profile_enable(True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): m.forward_features(feats)
b.record(); torch.cuda.synchronize()
r = ops.profile_read(); ops.profile_enable(False)
ms = a.elapsed_time(b) / 3
print(f"whole within-clip module, {clips} clips (T={T}, R50 641x641 pyramid): {ms:.2f} ms/forward = {clips / ms * 1e3:.0f} clips/s")
tot = sum(v["ms"] for v in r.values())
for k, v in sorted(r.items(), key=lambda kv: -kv[1]["ms"]):
    if v["timed"]:
        print(f"   {k:24s} {v['ms'] / 3:8.3f} ms  {100 * v['ms'] / tot:5.1f} %")
