"""What do the positional loads cost the q|k|v kernel?  Times qkv_pair_kernel with pos = None / one shared table / a per-clip tensor."""
import sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import ops, synth
clips = int(sys.argv[1]) if len(sys.argv) > 1 else 42
p = {k: v.cuda() for k, v in synth.axial_layer_params(0).items()}
pk = ops.pack_layer(p)
for hw in (21, 41):
    rows = clips * 2 * hw * hw
    x = torch.randn(rows, 256, device="cuda")
    tab = torch.randn(1, 2, hw, hw, 256, device="cuda")
    full = tab.expand(clips, -1, -1, -1, -1).contiguous()
    for name, pos in (("none", None), ("shared", tab), ("per-clip", full)):
        for _ in range(3):
            ops.traj_attn_fwd(x, x, x, pos, x, pk.attn_h, clips, 2, hw, hw, ops.AXIS_H)
        torch.cuda.synchronize()
        ops.profile_enable(True)
        for _ in range(5):
            ops.traj_attn_fwd(x, x, x, pos, x, pk.attn_h, clips, 2, hw, hw, ops.AXIS_H)
        torch.cuda.synchronize()
        r = ops.profile_read(); ops.profile_enable(False)
        print(f"{hw}x{hw} pos={name:9s}", {k: round(v["ms"] / max(v["timed"], 1) * 1e3, 1) for k, v in r.items() if v["timed"]})
