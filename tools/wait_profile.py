"""Debug: run one fused kernel on a libaxvs build with -DAXVS_WAIT_PROFILE and print the wait-cycle breakdown.
usage: AXVS_LIB=axial_vs_b200/libaxvs_prof.so python tools/wait_profile.py [ffn|traj] [clips]"""
import ctypes, os, sys, torch
sys.path.insert(0, ".")
from axial_vs_b200 import _lib, ops, synth
which = sys.argv[1] if len(sys.argv) > 1 else "ffn"
clips = int(sys.argv[2]) if len(sys.argv) > 2 else 32
hw = int(sys.argv[3]) if len(sys.argv) > 3 else 41
lib = _lib.load()
rows = clips * 2 * hw * hw
p = {k: v.cuda() for k, v in synth.axial_layer_params(0).items()}
pk = ops.pack_layer(p)
x = torch.randn(rows, 256, device="cuda")
buf = (ctypes.c_ulonglong * 64)()
def run():
    if which == "ffn":
        return ops.ln_ffn_fwd(x, pk)
    if which == "qkvd":
        ops.set_fusion(4)
        return ops.traj_attn_fwd(x, x, x, x, x, pk.attn_h, clips, 2, 41, 41, ops.AXIS_H)
    if which == "attn":
        ops.set_fusion(4)
        return ops.traj_attn_fwd(x, x, x, x, x, pk.attn_h, clips, 2, hw, hw, ops.AXIS_H)
    if which == "qkv":
        ops.set_fusion(3)
        return ops.traj_attn_fwd(x, x, x, None, x, pk.attn_h, clips, 2, 41, 41, ops.AXIS_H)
    return ops.traj_attn_fwd(x, x, x, None, x, pk.attn_h, clips, 2, 41, 41, ops.AXIS_H)
for _ in range(2): run()
lib.axvs_debug_read_waits(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
lib.axvs_debug_read_waits(buf)
v = list(buf)
tiles = (rows + 127) // 128
ctas = min(tiles, int(os.environ.get('AXVS_DEBUG_SMS', 148)))
if "--pair-ctas" in sys.argv:      # pair kernels: counters are flushed by the leader CTA of each pair only
    ctas = min((tiles + 1) // 2, 74)
    tiles = (tiles + 1) // 2
print(f"{which}: rows {rows} tiles {tiles} time {e0.elapsed_time(e1):.3f} ms")
def show(name, base, labels):
    tot = v[base + len(labels)] / ctas
    print(f"  {name}: total {tot:,.0f} clk/CTA ({tot / (tiles / ctas):,.0f} per tile)")
    for i, l in enumerate(labels):
        print(f"      wait {l:10s} {v[base + i] / ctas:12,.0f}  ({100.0 * v[base + i] / max(v[base + len(labels)], 1):5.1f} %)")
if which == "ffn":
    show("MMA warp", 0, ["w_full", "-", "a_full", "h_ready", "acc_free"])
    show("epilogue g0 warp0", 8, ["s_full", "-", "acc_full", "bar.sync", "final:chunks", "final:all", "st_wait"])
    show("epilogue g1 warp4", 16, ["s_full", "-", "acc_full", "bar.sync", "final:chunks", "final:all", "st_wait"])
    show("W producer", 32, ["w_empty"])
elif which == "attn":
    tiles = clips * hw * 8          # work units (sequence, head)
    ctas = min(tiles, 148)
    show("attn_tc softmax g0 warp0", 52, ["s_full", "o_full"])
    show("attn_tc S issuer", 56, ["slot full", "buf_free", "-"])
    show("attn_tc PV issuer", 36, ["p_full"])
    show("attn_tc producer", 60, ["slot empty"])
elif which == "qkvd":
    show("qkv_direct MMA warp", 40, ["w_full", "s_empty", "a_full"])
    if "--pair-ctas" in sys.argv:
        show("qkv_pair epilogue warp 0", 24, ["s_full", "[tmem ld]", "[cvt + sts]", "[bulk issue]", "[bulk wait]", "[fence]"])
        show("qkv_pair A producer warp 0", 48, ["a_empty", "[cvt + sts]"])
    else:
        show("qkv_direct epilogue g0", 44, ["s_full"])
        show("qkv_direct epilogue g1", 46, ["s_full"])
        show("qkv_direct A producer warp 0", 48, ["a_empty", "-"])
elif which == "qkv":
    show("qkv MMA warp", 40, ["w_full", "s_empty", "a_full"])
    show("qkv epilogue g0", 44, ["s_full"])
    show("qkv epilogue g1", 46, ["s_full"])
else:
    show("MMA warp", 0, ["w_full", "s_empty", "a_full", "o_ready", "pj_free", "[frames span]", "[gemm1 span]"])
    show("epilogue g0 warp0", 8, ["q2_full", "s_full(fr)", "s_full(pj)", "[q2 drain]", "[chunk hold]", "[o st_wait]", "[o pack]"])
    show("epilogue g1 warp4", 16, ["q2_full", "s_full(fr)", "s_full(pj)", "[q2 drain]", "[chunk hold]", "[o st_wait]", "[o pack]"])
    show("W producer", 32, ["w_empty"])
