"""traj_pair_kernel (cta_group::2) against traj_ts_kernel: identical outputs (same arithmetic, same accumulation order) at even / odd tile
counts and T = 2 / 5, then per-kernel times of one axial layer at 42 clips with the pair mode off / on."""
import sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import ops, synth
from axial_vs_b200.modules import TemporalAxialTrajectoryAttentionLayer

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
p = synth.axial_layer_params(0)
layer = TemporalAxialTrajectoryAttentionLayer(256, 1024, 0.0, 0.0, "relu", 8).eval()
layer.load_state_dict(p)
layer = layer.cuda()
ok = True
for (clips, T, H, W) in ((1, 2, 41, 41), (3, 2, 21, 21), (2, 5, 15, 20), (7, 2, 41, 41), (1, 2, 5, 5), (42, 2, 21, 21)):
    torch.manual_seed(clips * 100 + H)
    src = torch.randn(clips * T, H * W, 256, device="cuda")
    pos = torch.randn(clips, T, H, W, 256, device="cuda")
    with torch.no_grad():
        ops.set_pair_mode(0)
        a = layer(src, pos)[0].clone()
        ops.set_pair_mode(mode)
        b = layer(src, pos)[0].clone()
        b2 = layer(src, pos)[0].clone()
    torch.cuda.synchronize()
    d = (a - b).abs().max().item()
    print(f"clips {clips} T {T} {H}x{W}: tiles {(clips * T * H * W + 127) // 128}: max |pair - single| = {d:.3e}, run-to-run {(b - b2).abs().max().item():.1e}, finite {bool(torch.isfinite(b).all())}", flush=True)
    ok = ok and d < 1e-5
print("IDENTICAL" if ok else "MISMATCH")
for m in (0, mode, 0, mode):
    ops.set_pair_mode(m)
    clips, H, W = 42, 41, 41
    src = torch.randn(clips * 2, H * W, 256, device="cuda")
    pos = torch.randn(clips, 2, H, W, 256, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            layer(src, pos)
        torch.cuda.synchronize()
        ops.profile_enable(True)
        for _ in range(5):
            layer(src, pos)
        torch.cuda.synchronize()
        r = ops.profile_read()
        ops.profile_enable(False)
    print(f"pair mode {m}: " + "  ".join(f"{k.replace('_kernel', '')} {v['ms'] / v['timed'] * 1e3:.1f}" for k, v in r.items() if v["timed"]), flush=True)
ops.set_pair_mode(0)
