"""Per-kernel device time of one axial layer at each pyramid level of the bench workload (32 clips, T=2)."""
import sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import ops, synth
from axial_vs_b200.modules import TemporalAxialTrajectoryAttentionLayer

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 32
p = synth.axial_layer_params(0)
layer = TemporalAxialTrajectoryAttentionLayer(256, 1024, 0.0, 0.0, "relu", 8).eval()
layer.load_state_dict(p)
layer = layer.cuda()
shapes = ((21, 21), (41, 41))
if len(sys.argv) > 2:                      # e.g. "25x43,49x85": the res5 / res4 maps of the shipped 769 x 1345 VIPSeg configs
    shapes = tuple(tuple(int(v) for v in s_.split("x")) for s_ in sys.argv[2].split(","))
for (H, W) in shapes:
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
    rows = clips * 2 * H * W
    print(f"level {H}x{W}: rows {rows}, tiles {(rows + 127) // 128}")
    for k, v in r.items():
        if v["timed"]:
            print(f"   {k:26s} {v['ms'] / v['timed'] * 1e3:8.1f} us/launch  x{v['timed'] // 5}/layer   {v['flops'] / v['ms'] / 1e9 if v['ms'] else 0:7.1f} TF  {v['bytes'] / v['ms'] / 1e6 if v['ms'] else 0:7.0f} GB/s")
