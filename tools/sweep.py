"""Axial-trajectory attention micro-benchmark sweep (BASELINE configs[4]): T in {2,5,10}, H=W in {41,81,161}, C=256, heads=8.
One axial layer per point, batch chosen to give >= 4 waves of tiles; prints per-kernel TFLOP/s and the layer's fraction of the
measured sustained bf16 peak.  usage: python tools/sweep.py [--json out.json]"""
import json, sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import ops, synth
from axial_vs_b200.modules import TemporalAxialTrajectoryAttentionLayer

peak = 1381.5
try:
    peak = json.load(open("MEASURED_PEAKS.json")).get("bf16_tflops_sustained", peak)
except Exception:
    pass
layer = TemporalAxialTrajectoryAttentionLayer(256, 1024, 0.0, 0.0, "relu", 8).eval()
layer.load_state_dict(synth.axial_layer_params(0))
layer = layer.cuda()
rows_out = []
for T in (2, 5, 10):
    for HW in (41, 81, 161):
        tokens_per_clip = T * HW * HW
        B = max(1, min(64, (4 * 148 * 128 + tokens_per_clip - 1) // tokens_per_clip))
        src = torch.randn(B * T, HW * HW, 256, device="cuda")
        pos = torch.randn(B, T, HW, HW, 256, device="cuda")
        with torch.no_grad():
            for _ in range(2):
                layer(src, pos)
            torch.cuda.synchronize()
            ops.profile_enable(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                layer(src, pos)
            e1.record()
            torch.cuda.synchronize()
            prof = ops.profile_read()
            ops.profile_enable(False)
        ms = e0.elapsed_time(e1) / 3
        fl = B * T * HW * HW * 256 * (20 * 256 + 8 * T * 256 + 4 * T * (HW + HW) + 8 * T + 4 * 1024)   # SURVEY.md section 8d / BASELINE.md section 3
        tf = fl / (ms * 1e-3) / 1e12
        ks = {k: round(v["flops"] / v["ms"] / 1e9, 1) for k, v in prof.items() if v["timed"] and v["ms"] > 0}
        sh = {k: round(v["ms"] / sum(x["ms"] for x in prof.values()), 3) for k, v in prof.items() if v["timed"]}
        rows_out.append(dict(T=T, HW=HW, clips=B, ms=round(ms, 3), tflops=round(tf, 1), frac=round(tf / peak, 3), kernel_tflops=ks, kernel_share=sh))
        print(f"T={T:2d} {HW:3d}x{HW:<3d} clips={B:2d}: {ms:8.3f} ms  {tf:6.1f} TFLOP/s = {100 * tf / peak:4.1f} % of sustained peak | " +
              " ".join(f"{k.replace('_kernel', '')}:{v}" for k, v in ks.items()))
        del src, pos
        torch.cuda.empty_cache()
if "--json" in sys.argv:
    json.dump(rows_out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
