import sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import ops, synth
clips = int(sys.argv[1]) if len(sys.argv) > 1 else 6
p = {k: v.cuda() for k, v in synth.axial_layer_params(0).items()}
pk = ops.pack_layer(p)
rows = clips * 2 * 41 * 41
x = torch.randn(rows, 256, device="cuda")
out = ops.traj_attn_fwd(x, x, x, None, x, pk.attn_h, clips, 2, 41, 41, ops.AXIS_H)
torch.cuda.synchronize()
print("ok", rows, float(out.abs().mean()))
