"""Debug: parity error and time of the kMaX AxialAttention2D drop-in (row f3) at the R50 641x641 shapes (84 frames, 512 channels).
usage: python tools/debug/bench_kmax.py [simt]    (simt: the fp32 SIMT attention core everywhere instead of mma.sync for 33..48 positions)"""
import sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import synth, ops
from axial_vs_b200.kmax_axial import AxialAttention2D
from oracle import kmax_oracle as KO          # checker only (debug tool)
from axial_vs_b200 import _lib
_lib.load().axvs_set_kmax_tensor_cores(0 if 'simt' in sys.argv[1:] else 1)

for N, C, H, W in [(84, 512, 21, 21), (84, 512, 41, 41)]:
    ph, pw = synth.kmax_axial_params(1, C), synth.kmax_axial_params(2, 1024)
    m = AxialAttention2D(C, query_shape=[H, W]).eval()
    m._height_axis.load_state_dict(ph); m._width_axis.load_state_dict(pw); m.cuda()
    x = synth.randn(3, N, C, H, W).cuda()
    with torch.no_grad():
        y = m(x)
        ref = KO.axial_attention_2d(x[:2].cpu(), ph, pw)
        err = ((y[:2].cpu() - ref).abs().max() / ref.abs().max()).item()
        for _ in range(2): m(x)
        torch.cuda.synchronize()
        ops.profile_enable(True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): m(x)
        b.record(); torch.cuda.synchronize()
        r = ops.profile_read(); ops.profile_enable(False)
    print(f"AxialAttention2D {N}x{C}x{H}x{W}: max-normalised error {err:.2e}; {a.elapsed_time(b) / 5:.3f} ms/forward")
    for k, v in r.items():
        if v["timed"]:
            print(f"   {k:26s} {v['ms'] / 5 * 1e3:9.1f} us  ({v['timed'] // 5} launches)  {v.get('tflops', 0):.1f} TFLOP/s")
