import sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import ops, synth
p = {k: v.cuda() for k, v in synth.axial_layer_params(0).items()}
pk = ops.pack_layer(p)
x = torch.randn(42 * 2 * 41 * 41, 256, device="cuda")
for m in (0, 8, 1, 0, 8, 1):
    ops.set_pair_mode(m)
    for _ in range(3): ops.ln_ffn_fwd(x, pk)
    torch.cuda.synchronize()
    ops.profile_enable(True)
    for _ in range(5): ops.ln_ffn_fwd(x, pk)
    torch.cuda.synchronize()
    r = ops.profile_read(); ops.profile_enable(False)
    print(m, {k: round(v["ms"] / v["timed"] * 1e3, 1) for k, v in r.items() if v["timed"]})
