"""Debug: timeline of ONE tile of traj_ts_kernel on CTA 0 (profile build: AXVS_LIB=axial_vs_b200/libaxvs_prof.so)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import _lib, ops, synth
clips = int(sys.argv[1]) if len(sys.argv) > 1 else 42
import os
if os.environ.get('AXVS_PAIR'): print('pair mode', os.environ['AXVS_PAIR'], '(E0/E1 = leader CTA groups 0/1, E2/E3 = peer CTA)')
ln = len(sys.argv) > 2 and sys.argv[2] == "w"
lib = _lib.load()
rows = clips * 2 * 41 * 41
p = {k: v.cuda() for k, v in synth.axial_layer_params(0).items()}
pk = ops.pack_layer(p)
x = torch.randn(rows, 256, device="cuda")
buf = (ctypes.c_ulonglong * 512)()
for _ in range(3):
    ops.traj_attn_fwd(x, x, x, None, x, pk.attn_h, clips, 2, 41, 41, ops.AXIS_H)
lib.axvs_debug_read_trace(buf)
ops.traj_attn_fwd(x, x, x, None, x, pk.attn_h, clips, 2, 41, 41, ops.AXIS_H)
lib.axvs_debug_read_trace(buf)
v = list(buf)
names = {0: "I tile start", 3: "I xd copied", 1: "I gemm1 stages free", 2: "I gemm1 issued", 50: "I wait o_ready", 51: "I o_ready", 52: "I pj stages free", 53: "I gemm3 issued"}
for f in range(2):
    names[60 + 2 * f] = f"I copy x_{f} start"; names[61 + 2 * f] = f"I copy x_{f} issued"
    for ci in range(4):
        k = 4 * f + ci
        names[10 + 4 * k] = f"I chunk {k} (stage {ci & 1}) wait stage"; names[11 + 4 * k] = f"I chunk {k} stage free"
        names[12 + 4 * k] = f"I chunk {k} unit0 issued"; names[13 + 4 * k] = f"I chunk {k} unit1 issued"
for g in range(4):
    b = 100 + 100 * g
    names[b] = f"E{g} wait q2"; names[b + 1] = f"E{g} q2 full"; names[b + 2] = f"E{g} q2 drained/released"
    for j in range(4):
        names[b + 10 + 4 * j] = f"E{g} chunk {j} wait"; names[b + 11 + 4 * j] = f"E{g} chunk {j} full"; names[b + 12 + 4 * j] = f"E{g} chunk {j} released"
        names[b + 13 + 4 * j] = f"E{g} chunk {j} o updated"
    names[b + 40] = f"E{g} o-final start"; names[b + 41] = f"E{g} o_ready arrived"; names[b + 42] = f"E{g} resid prefetched, wait pj"
    names[b + 43] = f"E{g} pj full"; names[b + 44] = f"E{g} pj stage released"; names[b + 45] = f"E{g} output done"
print(f'CTA 0: {v[502] - v[500]} ns, {v[503] - v[501]} clk -> {(v[503] - v[501]) / max(1, v[502] - v[500]):.3f} GHz')
ev = sorted((t, i) for i, t in enumerate(v[:500]) if t)
t0 = ev[0][0]
prev = t0
for t, i in ev:
    print(f"{t - t0:8d}  (+{t - prev:6d})  {names.get(i, i)}")
    prev = t
