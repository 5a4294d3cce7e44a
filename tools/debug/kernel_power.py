"""Board power and SM clock while ONE part of the hot path runs in a loop (NVML samples every 20 ms over ~1.5 s each): which kernels push the board
to its power limit?  Parts: the three kernels of one TrajectoryAttention (q|k|v, attention, temporal), the FFN kernel alone, the whole axial layer."""
import sys, time, threading, torch
sys.path.insert(0, ".")
import pynvml
from axial_vs_b200 import ops, synth
from axial_vs_b200.modules import TemporalAxialTrajectoryAttentionLayer
clips, H, W = 42, 41, 41
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
p = synth.axial_layer_params(0)
pk = ops.pack_layer({k: v.cuda() for k, v in p.items()})
layer = TemporalAxialTrajectoryAttentionLayer(256, 1024, 0.0, 0.0, "relu", 8).eval()
layer.load_state_dict(p); layer.cuda()
rows = clips * 2 * H * W
x = torch.randn(rows, 256, device="cuda")
src = x.view(clips * 2, H * W, 256)
tab = torch.randn(1, 2, H, W, 256, device="cuda")
parts = {
    "q|k|v + attention + temporal (H pass)": lambda: ops.traj_attn_fwd(x, x, x, tab, x, pk.attn_h, clips, 2, H, W, ops.AXIS_H),
    "FFN kernel (LayerNorm1 + FFN + LayerNorm2)": lambda: ops.ln_ffn_fwd(x, pk),
    "whole axial layer": lambda: layer(src, tab.expand(clips, -1, -1, -1, -1)),
}
for name, fn in parts.items():
    with torch.no_grad():
        for _ in range(3): fn()
        torch.cuda.synchronize()
        time.sleep(1.0)                                   # let the board cool down to idle power between parts
        samples, stop = [], False
        def sampler():
            while not stop:
                samples.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                time.sleep(0.02)
        t = threading.Thread(target=sampler); t.start()
        t0 = time.time(); n = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < 1.5:
            for _ in range(20): fn()
            n += 20
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        stop = True; t.join()
    late = samples[len(samples) // 2:]
    print(f"{name:46s} {e0.elapsed_time(e1) / n * 1e3:8.1f} us per call   power (second half) {sum(s[0] for s in late) / len(late):6.0f} W   "
          f"SM clock {sum(s[1] for s in late) / len(late):6.0f} MHz   (first samples: {[round(s[0]) for s in samples[:4]]} W)")
