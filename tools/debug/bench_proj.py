"""Time the within-clip input / output projections at the R50 pyramid sizes of the bench workload (42 clips x 2 frames)."""
import sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import synth
from axial_vs_b200.projections import InputProjection, OutputProjection
n = 84
for c, HW in ((2048, 21), (1024, 41), (512, 81)):
    pin, pout = synth.proj_params(1, c)
    mi, mo = InputProjection(c).eval(), OutputProjection(c).eval()
    mi.load_state_dict(pin); mi.cuda(); mo.cuda()
    x = torch.randn(n, c, HW, HW, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            tok = mi(x); y = mo(tok, HW, HW)
        a, b, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        for _ in range(10): tok = mi(x)
        b.record()
        for _ in range(10): y = mo(tok, HW, HW)
        c2.record(); torch.cuda.synchronize()
    px = n * HW * HW
    t_in, t_out = a.elapsed_time(b) / 10, b.elapsed_time(c2) / 10
    print(f"c={c:4d} {HW}x{HW}: input_proj {t_in*1e3:7.1f} us ({2*px*c*256/t_in/1e9:6.1f} TFLOP/s, {px*(c*4+1024*3)/t_in/1e6:6.0f} GB/s)   "
          f"output_proj {t_out*1e3:7.1f} us ({2*px*c*256/t_out/1e9:6.1f} TFLOP/s, {px*(1024+c*4*3)/t_out/1e6:6.0f} GB/s)")
from axial_vs_b200 import ops
for c, HW in ((2048, 21), (512, 81)):
    pin, pout = synth.proj_params(1, c)
    mi, mo = InputProjection(c).eval().cuda(), OutputProjection(c).eval().cuda()
    x = torch.randn(n, c, HW, HW, device="cuda")
    with torch.no_grad():
        tok = mi(x); y = mo(tok, HW, HW); torch.cuda.synchronize()
        ops.profile_enable(True)
        tok = mi(x); torch.cuda.synchronize(); r1 = ops.profile_read(); ops.profile_enable(False)
        ops.profile_enable(True)
        y = mo(tok, HW, HW); torch.cuda.synchronize(); r2 = ops.profile_read(); ops.profile_enable(False)
    print(c, HW, "input:", {k: round(v["ms"] * 1e3, 1) for k, v in r1.items() if v["timed"]}, "output:", {k: round(v["ms"] * 1e3, 1) for k, v in r2.items() if v["timed"]})
