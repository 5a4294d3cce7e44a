import sys, torch
sys.path.insert(0, '.')
from axial_vs_b200 import ops, synth
from oracle import traj_oracle as O
Bp, F, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
p = {}
synth.traj_attn_params(torch.Generator().manual_seed(11), "", 256, p)
q = synth.randn(111, Bp, F * n, 256); v = synth.randn(211, Bp, F * n, 256)
pk = ops.pack_ta({k_: t.cuda() for k_, t in p.items()})
ref, _ = O.trajectory_attention(q, q, v, p, F)
qc, vc = q.reshape(-1, 256).cuda(), v.reshape(-1, 256).cuda()
for level in (0, 1):
    ops.set_fusion(level)
    out = ops.traj_attn_fwd(qc, qc, vc, None, None, pk, Bp, F, n, 1, ops.AXIS_NONE)
    torch.cuda.synchronize()
    e = ((out.cpu() - ref.reshape(-1, 256)).abs().max() / ref.abs().max()).item()
    print("level", level, "err", e, flush=True)
