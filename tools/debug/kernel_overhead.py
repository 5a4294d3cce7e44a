import ctypes, sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import _lib, ops, synth
lib = _lib.load()
clips=42
rows = clips * 2 * 41 * 41
p = {k: v.cuda() for k, v in synth.axial_layer_params(0).items()}
pk = ops.pack_layer(p)
x = torch.randn(rows, 256, device="cuda")
buf = (ctypes.c_ulonglong * 512)()
wb = (ctypes.c_ulonglong * 64)()
def run(): return ops.traj_attn_fwd(x, x, x, None, x, pk.attn_h, clips, 2, 41, 41, ops.AXIS_H)
for _ in range(3): run()
lib.axvs_debug_read_trace(buf); lib.axvs_debug_read_waits(wb)
ops.profile_enable(True)
run()
torch.cuda.synchronize()
r = ops.profile_read(); ops.profile_enable(False)
lib.axvs_debug_read_trace(buf); lib.axvs_debug_read_waits(wb)
v = list(buf); w = list(wb)
print("event-timed:", {k: round(vv["ms"]*1e3,1) for k, vv in r.items() if vv["timed"]})
print(f"traj CTA 0 lifetime: {v[502]-v[500]} ns, {v[503]-v[501]} clk -> {(v[503]-v[501])/max(1,v[502]-v[500]):.3f} GHz")
print(f"traj issuer loop total (avg over CTAs): {w[7]/148:.0f} clk")
