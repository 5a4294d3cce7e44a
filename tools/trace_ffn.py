"""Debug: timeline of ONE tile of ffn_n256_pair_kernel on the leader CTA of pair 0 (profile build: AXVS_LIB=axial_vs_b200/libaxvs_prof.so)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import _lib, ops, synth
lib = _lib.load()
p = {k: v.cuda() for k, v in synth.axial_layer_params(0).items()}
pk = ops.pack_layer(p)
x = torch.randn(42 * 2 * 41 * 41, 256, device="cuda")
buf = (ctypes.c_ulonglong * 512)()
for _ in range(3): ops.ln_ffn_fwd(x, pk)
lib.axvs_debug_read_trace(buf)
ops.ln_ffn_fwd(x, pk)
lib.axvs_debug_read_trace(buf)
v = list(buf)
names = {0: "I tile start", 1: "I gemm1(0) issued"}
for j in range(4):
    names[10 + 4 * j] = f"I wait h({j})"; names[11 + 4 * j] = f"I h({j}) ready"; names[12 + 4 * j] = f"I gemm2({j}) issued"; names[13 + 4 * j] = f"I gemm1({j + 1}) issued"
for g in range(2):
    b = 100 + 100 * g
    for j in range(4):
        names[b + 4 * j] = f"E{g} wait chunk {j}"; names[b + 4 * j + 1] = f"E{g} chunk {j} full"; names[b + 4 * j + 2] = f"E{g} h({j}) written"
    names[b + 40] = f"E{g} resid loaded, wait acc"; names[b + 41] = f"E{g} acc full"; names[b + 42] = f"E{g} acc released"; names[b + 43] = f"E{g} tile done"
ev = sorted((t, i) for i, t in enumerate(v[:500]) if t)
t0 = ev[0][0]; prev = t0
for t, i in ev:
    print(f"{t - t0:8d}  (+{t - prev:6d})  {names.get(i, i)}"); prev = t
