"""Per-kernel times of ONE TrajectoryAttention along H and along W at a given map size (the attention kernel's cost depends on the frame length n)."""
import sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import ops, synth
clips = int(sys.argv[1]) if len(sys.argv) > 1 else 30
H, W = (int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "49x85").split("x"))
p = {k: v.cuda() for k, v in synth.axial_layer_params(0).items()}
pk = ops.pack_layer(p)
rows = clips * 2 * H * W
x = torch.randn(rows, 256, device="cuda")
tab = torch.randn(1, 2, H, W, 256, device="cuda")
for name, axis, w in (("H pass (n = %d)" % H, ops.AXIS_H, pk.attn_h), ("W pass (n = %d)" % W, ops.AXIS_W, pk.attn_w)):
    for _ in range(3):
        ops.traj_attn_fwd(x, x, x, tab, x, w, clips, 2, H, W, axis)
    torch.cuda.synchronize()
    ops.profile_enable(True)
    for _ in range(5):
        ops.traj_attn_fwd(x, x, x, tab, x, w, clips, 2, H, W, axis)
    torch.cuda.synchronize()
    r = ops.profile_read(); ops.profile_enable(False)
    print(f"{H}x{W} {name:18s}", {k: round(v["ms"] / max(v["timed"], 1) * 1e3, 1) for k, v in r.items() if v["timed"]})
