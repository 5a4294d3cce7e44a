"""Does keeping the intermediates (q|k|v, x_f / x_diag images, LN image) L2-resident pay?  One axial layer over 42 clips run as ONE call vs
as a loop over chunks of c clips that reuse the same workspace (graph-replayed, so launch overhead is excluded)."""
import sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import ops, synth
from axial_vs_b200.modules import TemporalAxialTrajectoryAttentionLayer

clips = 42
p = synth.axial_layer_params(0)
layer = TemporalAxialTrajectoryAttentionLayer(256, 1024, 0.0, 0.0, "relu", 8).eval()
layer.load_state_dict(p)
layer = layer.cuda()
for (H, W) in ((41, 41), (21, 21)):
    src = [torch.randn(clips * 2, H * W, 256, device="cuda") for _ in range(2)]
    pos = torch.randn(1, 2, H, W, 256, device="cuda").expand(clips, 2, H, W, 256)
    for c in (42, 21, 14, 7, 6, 4, 3, 2, 1):
        def run(s):
            outs = []
            for i in range(0, clips, c):
                outs.append(layer(s[2 * i:2 * (i + c)], pos[i:i + c])[0])
            return outs
        with torch.no_grad():
            for _ in range(2):
                run(src[0])
            torch.cuda.synchronize()
            gs = []
            for s in src:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    run(s)
                gs.append(g)
            for g in gs: g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                for g in gs: g.replay()
            e1.record()
            torch.cuda.synchronize()
        print(f"{H}x{W}: chunks of {c:2d} clips ({(c * 2 * H * W + 127) // 128:4d} tiles): {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us per 42-clip layer", flush=True)
        del gs
