import sys, torch
sys.path.insert(0, '.')
from axial_vs_b200 import ops, synth
from oracle import traj_oracle as O
rows = int(sys.argv[1])
p = synth.axial_layer_params(3)
pk = ops.pack_layer({k: v.cuda() for k, v in p.items()})
x = synth.randn(5, rows, 256)
ref = O._ffn_tail(x, p)
for pair in (0, 1):
    ops.set_pair_mode(pair)
    out = ops.ln_ffn_fwd(x.cuda(), pk)
    torch.cuda.synchronize()
    print("rows", rows, "pair", pair, "err", ((out.cpu() - ref).abs().max() / ref.abs().max()).item(), flush=True)
