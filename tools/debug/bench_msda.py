"""Time the MSDeformAttn spatial encoder layer at the R50 641x641 pyramid (42 clips x 2 frames, 8683 tokens per frame)."""
import sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import synth, ops, msda
n = int(sys.argv[1]) if len(sys.argv) > 1 else 84
shapes = [(21, 21), (41, 41), (81, 81)]
Len = sum(h * w for h, w in shapes)
m = msda.MSDeformAttnTransformerEncoderLayer(256, 1024, 0.0, "relu", 3, 8, 4).eval()
m.load_state_dict(synth.msda_layer_params(1)); m.cuda()
src = torch.randn(n, Len, 256, device="cuda"); pos = torch.randn(n, Len, 256, device="cuda")
ref = msda.reference_points(shapes, n, "cuda")
with torch.no_grad():
    for _ in range(2): m(src, pos, ref, shapes)
    torch.cuda.synchronize()
    ops.profile_enable(True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): m(src, pos, ref, shapes)
    b.record(); torch.cuda.synchronize()
    r = ops.profile_read(); ops.profile_enable(False)
rows = n * Len
ms = a.elapsed_time(b) / 3
fl = rows * (2 * 256 * (256 + 288 + 256) + 4 * 256 * 1024)
print(f"msda layer: {rows} tokens, {ms:.3f} ms/layer, {fl / ms / 1e9:.1f} TFLOP/s (linear layers only)")
for k, v in r.items():
    if v["timed"]:
        print(f"   {k:24s} {v['ms'] / 3 * 1e3:9.1f} us/layer  ({v['timed'] // 3} launches)")
