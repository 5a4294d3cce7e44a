#!/usr/bin/env python
"""bench.py -- clips/s of the within-clip tracking module's temporal hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clips C] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Video-kMaX R50 + MaXTron within-clip tracking on VIPSeg-shaped synthetic clips,
T=2 frames of 641x641 -> temporal levels res5 (21x21) and res4 (41x41), C=256, 2 stages x TemporalEncoder(2 axial-
trajectory layers), random-init weights (xavier), synthetic N(0,1) features.  One step = that hot path over one
batch of `--clips` clips per GPU (weak scaling: every rank processes its own clips, no collective in the data path;
a small per-clip output summary is all-gathered once per step).  Algorithmic work: 60.5 GFLOP per clip.

`--impl reference` times the reference's CPU implementation of the same path on the host cores: the UNMODIFIED reference modules
from oracle/_ref/ (placed there by oracle/build_ref.py; `kind: "reference"`), or the oracle port (`kind: "port"`) when they are absent.

`--config` selects the other BASELINE.json configurations (default 1 = the metric's configuration):
    3  cross-clip tracking module on one 64-clip video, Q = 128, 4 layers, clip-sharded over the GPUs (strong scaling, NCCL all-gathers)
    4  Tube-Link flavour: T = 5, 480x640 -> temporal levels 15x20 + 30x40, 6 encoder layers x 1 temporal layer, gamma skip
    5  axial-trajectory micro-benchmark sweep: T in {2,5,10} x H=W in {41,81,161}, one layer per point
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_FRAMES = 2
LEVELS = [(21, 21), (41, 41)]       # res5, res4 of a 641x641 frame (low -> high resolution, WC/msdeformattn.py:413)
STAGES = 2                           # NUM_STAGES
LAYERS_PER_STAGE = 2                 # TEMPORAL_LAYERS // NUM_STAGES  (maxtron_wc_r50.yaml)
METRIC = "video clips/sec"


def flops_per_clip() -> float:
    C, dffn, T = 256, 1024, T_FRAMES
    tot = 0.0
    for (H, W) in LEVELS:
        tokens = T * H * W
        tot += tokens * C * (20 * C + 8 * T * C + 4 * T * (H + W) + 8 * T + 4 * dffn)
    return tot * STAGES * LAYERS_PER_STAGE


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], source="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def reference_encoders(device="cpu", dtype=torch.float32):
    """The hot path's modules for the baseline legs: (stages, kind).  kind "reference" = the unmodified reference TemporalEncoder
    from oracle/_ref (same state-dict keys, so the synthetic weights load strictly); kind "port" = the oracle restatement."""
    from axial_vs_b200 import synth
    from oracle import build_ref
    ref = build_ref.load()
    if ref is not None:
        ta, _ = ref
        stages = []
        for s_ in range(STAGES):
            enc = ta.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "axial-trajectory", LAYERS_PER_STAGE).eval()
            enc.load_state_dict(synth.encoder_params(s_, LAYERS_PER_STAGE), strict=True)
            stages.append(enc.to(device=device, dtype=dtype))
        return stages, "reference"
    from oracle import traj_oracle as O
    return [O.split_encoder_params({k: v.to(device) for k, v in synth.encoder_params(s_, LAYERS_PER_STAGE).items()}) for s_ in range(STAGES)], "port"


def reference_clip_runner(clips: int = 1, device: str = "cpu"):
    """Returns (fn, threads, kind): fn() runs the hot path of `clips` clips through the reference's own modules (or the port)."""
    from axial_vs_b200 import synth
    from oracle import traj_oracle as O
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    stages, kind = reference_encoders(device)
    le = synth.level_embed(99)
    srcs = [synth.randn(10 + i, clips * T_FRAMES, H * W, 256).to(device) for i, (H, W) in enumerate(LEVELS)]
    poss = [O.level_pos3d(1, T_FRAMES, H, W, le[i]).expand(clips, -1, -1, -1, -1).contiguous().to(device) for i, (H, W) in enumerate(LEVELS)]

    @torch.no_grad()
    def run():
        cur = list(srcs)
        for st in stages:
            for i in range(len(LEVELS)):
                if kind == "reference":
                    cur[i] = st(cur[i], poss[i])[0]
                else:
                    cur[i], _, _ = O.temporal_encoder(cur[i], poss[i], st)
        return cur

    return run, torch.get_num_threads(), kind


def oracle_clip_runner():
    run, cores, _ = reference_clip_runner(1, "cpu")
    return run, cores


def cpu_baseline_sample(budget_s: float = 12.0):
    run, cores, kind = reference_clip_runner(1, "cpu")
    run()                                  # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        run()
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 200:
            break
    what = "the unmodified reference modules (oracle/_ref)" if kind == "reference" else "the oracle port of the reference modules"
    return {"value": n / el, "unit": "clips/s", "cores": cores, "kind": kind,
            "sample": f"{n} clips (T=2, res5 21x21 + res4 41x41, 2 stages x 2 layers) in {el:.1f} s, fp32 torch-CPU, {what}"}


def gpu_eager_baseline(clips: int, dev, steps: int = 5):
    """The reference's own eager-PyTorch path on THIS GPU (the bar the kernels have to beat): fp32, and under bf16
    autocast, same clips per step as the product arm, inputs resident.  A reported baseline, not on the product path."""
    out = {}
    try:
        run, _, kind = reference_clip_runner(clips, str(dev))
        for name, ctx in (("fp32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            def go():
                if ctx is None:
                    return run()
                with ctx:
                    return run()
            for _ in range(2):
                go()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                go()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"clips_per_s": round(clips / (ms * 1e-3), 1), "ms_per_step": round(ms, 3)}
        out["kind"] = kind
        out["what"] = f"{'unmodified reference modules' if kind == 'reference' else 'oracle port'} on cuda, eager, {clips} clips per step, {steps} steps"
    except Exception as e:                                  # e.g. out of memory: the reference materialises [B h, N, N] scores
        out["error"] = f"{type(e).__name__}: {e}"
    torch.cuda.empty_cache()
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run, cores, kind = reference_clip_runner(1, "cpu")
    for _ in range(max(1, min(args.warmup, 2))):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    el = time.perf_counter() - t0
    val = args.steps / el
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.clips),
            "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": kind,
                             "sample": "bounded sample of the workload: 1 clip per timed step (the CPU path is batch-size independent: clips run "
                                       "one after the other); " + ("the unmodified reference modules (oracle/_ref)" if kind == "reference" else
                                                                   "the oracle port of the reference modules (oracle/traj_oracle.py)") + ", all host threads"},
            "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(clips):
    return {"workload": "Video-kMaX R50 + MaXTron within-clip tracking hot path (BASELINE configs[1]): T=2, 641x641 -> "
                        "res5 21x21 + res4 41x41, C=256, 2 stages x TemporalEncoder(2 axial-trajectory layers)",
            "clips_per_gpu_per_step": clips, "gflop_per_clip": round(flops_per_clip() / 1e9, 2),
            "pos": "one PositionEmbeddingSine3D table per level shared by all clips (the reference's table is clip-independent)",
            "l2": "per-step activations exceed the 126 MB L2 and two input sets alternate"}


def bind_to_gpu_numa(gpu_index: int):
    """Pin this process to the CPUs NVML reports as affine to its GPU, so the pinned staging buffers are first-touched on the GPU's NUMA
    node.  Returns the CPU list as a short string (or why it was skipped): the bench line records it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{cpus[0]}-{cpus[-1]} ({len(cpus)} cpus)"
        return "no affine cpus reported"
    except Exception as e:
        return f"unchanged ({type(e).__name__})"


def pcie_probe(dev, world, dist, mb: int = 256, reps: int = 4):
    """Host<->device ceilings of THIS box at THIS rank count: every rank moves `mb` MiB of pinned memory at the same time (H2D alone,
    D2H alone, both directions at once on two streams); aggregate GB/s over all ranks, max-over-ranks time.  The e2e leg is bounded by
    the bidirectional figure (its H2D and D2H streams run concurrently)."""
    n = mb << 20
    h_a, h_b = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a, d_b = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d, d2h):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(dev)
        e0.record(cur)
        s1.wait_stream(cur); s2.wait_stream(cur)
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
        cur.wait_stream(s1); cur.wait_stream(s2)
        e1.record(cur)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return world * reps * n * (int(h2d) + int(d2h)) / (ms.item() * 1e-3) / 1e9

    run(True, True)                                          # warm-up (first touch of the pinned pages)
    return {"h2d_gbs": round(run(True, False), 1), "d2h_gbs": round(run(False, True), 1), "bidir_gbs": round(run(True, True), 1),
            "how": f"{mb} MiB pinned per rank and direction, {reps} copies, all {world} rank(s) at once, aggregate over ranks"}


def module_figure(dev, clip_counts=(16, 42), steps: int = 5):
    """Whole WithinClipTrackingModule.forward_features (input projections, positional tables, 2 x [MSDeformAttn spatial layer + 2
    axial-trajectory layers], output projections) at the R50 641x641 pyramid: clips/s with inputs resident, eager launches."""
    from axial_vs_b200 import synth, within_clip
    from types import SimpleNamespace
    chans, sizes = [2048, 1024, 512], [(21, 21), (41, 41), (81, 81)]
    shape = {f"res{5 - i}": SimpleNamespace(channels=chans[i], stride=2 ** (5 - i)) for i in range(3)}
    out = {}
    try:
        m = within_clip.WithinClipTrackingModule(
            shape, transformer_dropout=0.0, transformer_attn_drop=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
            transformer_num_stages=2, transformer_spatial_layers=2, transformer_temporal_layers=4,
            transformer_temporal_attn_type="axial-trajectory", conv_dims=256, transformer_spatial_in_features=["res3", "res4", "res5"],
            transformer_temporal_in_features=["res4", "res5"], num_clip_frames=2, cross_clip_training=True).eval()
        m.load_state_dict(synth.within_clip_module_params(7, chans, temporal_layers_per_stage=2), strict=True)
        m.to(dev)
        for clips in clip_counts:
            feats = {f"res{5 - i}": torch.randn(clips * 2, chans[i], *sizes[i], device=dev) for i in range(3)}
            with torch.no_grad():
                for _ in range(2):
                    m.forward_features(feats)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    m.forward_features(feats)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[f"clips_{clips}"] = {"clips_per_s": round(clips / (ms * 1e-3), 1), "ms_per_forward": round(ms, 3)}
            try:                                    # the same forward replayed from a CUDA graph (the module is allocation-free and never synchronises)
                with torch.no_grad():
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        m.forward_features(feats)
                    g.replay()
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(steps):
                        g.replay()
                    e1.record()
                    torch.cuda.synchronize()
                msg = e0.elapsed_time(e1) / steps
                out[f"clips_{clips}"].update({"graph_clips_per_s": round(clips / (msg * 1e-3), 1), "graph_ms_per_forward": round(msg, 3)})
                del g
            except Exception as e:
                out[f"clips_{clips}"]["graph_error"] = f"{type(e).__name__}: {str(e).splitlines()[0]}"
                torch.cuda.synchronize()
            del feats
            torch.cuda.empty_cache()
        out["what"] = "WithinClipTrackingModule.forward_features, R50 pyramid (res5 2048x21^2, res4 1024x41^2, res3 512x81^2), T=2, inputs resident; eager launches, and (graph_*) the same forward replayed from a CUDA graph"
    except Exception as e:
        out["error"] = f"{type(e).__name__}: {e}"
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clips", type=int, default=42, help="clips per GPU per step (42: res5 = 1.96 and res4 = 7.45 waves of 128-row tiles over 148 SMs)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="1 (default): replay each step's kernels from a CUDA graph (launch overhead removed); 0: eager launches")
    ap.add_argument("--per-step", type=int, default=0, help="debug: print per-step device times of the timed regions to stderr")
    ap.add_argument("--clip-chunks", type=int, default=1, help="split each level's clips into this many groups, one CUDA stream per (level, group)")
    ap.add_argument("--full-pos", type=int, default=0, help="1: materialise the positional table for every clip ([B,T,H,W,C], as the reference passes it) instead of sharing one table")
    ap.add_argument("--level-streams", type=int, default=1, help="1: run the two pyramid levels on two CUDA streams (default), 0: serially")
    ap.add_argument("--config", type=int, default=1, choices=[1, 3, 4, 5], help="BASELINE.json configuration (1 = the metric's; see the module docstring)")
    ap.add_argument("--host-dtype", default="fp32", choices=["fp32", "bf16"], help="dtype of the pinned host buffers of the e2e leg (fp32 = what the reference "
                    "passes between modules; bf16 = opt-in half-size host boundary, converted on the device)")
    ap.add_argument("--d2h", default="full", choices=["full", "summary"], help="e2e device->host read: the full output feature maps (default) or a per-clip "
                    "pooled summary [clips, 256] per level (what a caller that keeps the features on the GPU would read back)")
    ap.add_argument("--no-extras", action="store_true", help="skip the reported-only legs (whole-module figure, eager-GPU baseline of the reference, PCIe probe)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.config != 1:
        import bench_configs
        bench_configs.run(args)
        return

    import torch.distributed as dist
    from axial_vs_b200 import modules, ops, sharding, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: a CUDA device is required (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = bind_to_gpu_numa(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    clips = args.clips
    encoders = []
    for s in range(STAGES):
        enc = modules.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "axial-trajectory", LAYERS_PER_STAGE).eval()
        enc.load_state_dict(synth.encoder_params(s, LAYERS_PER_STAGE), strict=True)
        encoders.append(enc.to(dev))
    le = synth.level_embed(99).to(dev)
    pos = [ops.pos3d(1, T_FRAMES, H, W, le[i].contiguous(), dev).expand(clips, -1, -1, -1, -1) for i, (H, W) in enumerate(LEVELS)]   # one table per level, shared by the clips (as PositionEmbeddingSine3D.table)
    if args.full_pos:
        pos = [p.contiguous() for p in pos]
    # two alternating input sets (device-resident for `value`, pinned host copies for `e2e`)
    hdt = torch.bfloat16 if args.host_dtype == "bf16" else torch.float32
    host_in = [[torch.randn(clips * T_FRAMES, H * W, 256, generator=torch.Generator().manual_seed(1000 * k + 10 * rank + i)).to(hdt).pin_memory()
                for i, (H, W) in enumerate(LEVELS)] for k in range(2)]
    dev_in = [[t.to(dev).float() for t in hs] for hs in host_in]
    # per-step outputs read back by the e2e leg: the full feature maps, or a pooled per-clip summary per level
    out_shapes = [(clips * T_FRAMES, H * W, 256) for (H, W) in LEVELS] if args.d2h == "full" else [(clips, 256) for _ in LEVELS]

    from axial_vs_b200 import within_clip

    @torch.no_grad()
    def hot_path(srcs, concurrent=True):
        cur = list(srcs)
        for enc in encoders:                       # the same TemporalEncoder object serves both levels (WC/msdeformattn.py:261-263)
            if args.level_streams and concurrent:
                outs = within_clip.run_levels_concurrent(enc, cur, pos, args.clip_chunks)    # independent levels / clip groups: one stream each
                cur = [o[0] for o in outs]
            else:
                for i in range(len(LEVELS)):
                    cur[i], _, _ = enc(cur[i], pos[i])
        return cur

    def gather_summary(outs):
        if world > 1:                              # per-clip summary, all-gathered after the path (never inside it)
            summ = outs[1].view(clips, -1, 256).mean(1)
            sharding.gather_clip_outputs(summ, clips * world)

    def step_resident(k):
        outs = hot_path(dev_in[k & 1])
        gather_summary(outs)
        return outs

    graphs = {}

    def step_resident_graph(k):
        # the library is allocation-free and never synchronises, so a whole step (both input sets) is captured once and replayed
        g = graphs.get(k & 1)
        if g is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                graphs[("out", k & 1)] = hot_path(dev_in[k & 1])
            graphs[k & 1] = g
        g.replay()
        gather_summary(graphs[("out", k & 1)])
        return graphs[("out", k & 1)]

    # e2e: the same step through the public nn.Module API with HOST buffers.  Three streams pipeline it the way a serving
    # loop would: H2D of step k+1 and D2H of step k-1 overlap the kernels of step k (NSLOT-deep pinned host + device staging: with
    # two slots the kernels of step k+2 had to wait for the D2H of step k, which takes about as long as a step -- any slip stalled the
    # GPU; three slots decouple them); every step still moves its own inputs in and its own outputs out inside the timed region.
    NSLOT = 3
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    stage_in = [[torch.empty_like(t, device=dev) for t in host_in[0]] for _ in range(NSLOT)]
    host_out2 = [[torch.empty(sh, dtype=hdt).pin_memory() for sh in out_shapes] for _ in range(NSLOT)]

    def e2e_compute(staged):
        """staged host-dtype inputs -> outputs in the host dtype / read-back shape (conversions and pooling run on the device)."""
        outs = hot_path([t.float() for t in staged] if hdt != torch.float32 else staged)
        if args.d2h == "summary":
            outs = [o.view(clips, -1, 256).mean(1) for o in outs]
        return [o.to(hdt) for o in outs] if hdt != torch.float32 else outs
    ev_in = [torch.cuda.Event() for _ in range(NSLOT)]        # H2D of slot finished
    ev_used = [torch.cuda.Event() for _ in range(NSLOT)]      # compute finished reading slot's staging inputs
    ev_comp = [torch.cuda.Event() for _ in range(NSLOT)]      # compute of slot finished
    ev_out = [torch.cuda.Event() for _ in range(NSLOT)]       # D2H of slot finished
    e2e_state = {"n": 0, "keep": [None] * NSLOT}
    e2e_graphs = {}

    def step_e2e(k):
        slot = e2e_state["n"] % NSLOT
        first = e2e_state["n"] < NSLOT
        e2e_state["n"] += 1
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(s_in):
            if not first:
                s_in.wait_event(ev_used[slot])
            for h, d_ in zip(host_in[k & 1], stage_in[slot]):
                d_.copy_(h, non_blocking=True)
            ev_in[slot].record(s_in)
        cur.wait_event(ev_in[slot])
        if args.graph:
            # the kernels of the step replay from a CUDA graph captured on this slot's staging buffers; its output buffers are
            # rewritten NSLOT steps later, after the D2H of this slot has been waited for
            if not first:
                cur.wait_event(ev_out[slot])
            g = e2e_graphs.get(slot)
            if g is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    e2e_graphs[("out", slot)] = e2e_compute(stage_in[slot])
                e2e_graphs[slot] = g
            g.replay()
            outs = e2e_graphs[("out", slot)]
        else:
            outs = e2e_compute(stage_in[slot])
        ev_used[slot].record(cur)
        if args.d2h == "full":
            gather_summary(outs)
        elif world > 1:
            sharding.gather_clip_outputs(outs[1].float(), clips * world)
        ev_comp[slot].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_comp[slot])
            if not first:
                pass                                   # host_out2[slot] was drained NSLOT steps ago (same stream, in order)
            for o, h in zip(outs, host_out2[slot]):
                h.copy_(o, non_blocking=True)
                if not args.graph:
                    o.record_stream(s_out)
            ev_out[slot].record(s_out)
        e2e_state["keep"][slot] = outs
        return outs

    def e2e_drain():
        torch.cuda.current_stream(dev).wait_stream(s_out)
        torch.cuda.current_stream(dev).wait_stream(s_in)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, drain=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = []
        e0.record()
        for k in range(steps):
            fn(k)
            if args.per_step:                          # debug: per-step device times (stderr), e.g. to see a power cap set in
                marks.append(torch.cuda.Event(enable_timing=True))
                marks[-1].record()
        if drain is not None:
            drain()                                    # the timed region ends when the last D2H has landed
        e1.record()
        barrier()
        if marks and rank == 0:
            ts = [e0.elapsed_time(m) for m in marks]
            print("[bench] per-step ms: " + " ".join(f"{b - a_:.2f}" for a_, b in zip([0.0] + ts[:-1], ts)), file=sys.stderr)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # the clock sampler is started BEFORE the warm-up: nvidia-smi / NVML initialisation takes driver locks for tens of
    # milliseconds, which must not fall inside the timed regions (it did: 10x outliers of the e2e figure)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.6)
    # kernels launched per step: counted by the library on one eager step (graph replays do not pass through its entry points)
    step_resident(0)
    torch.cuda.synchronize()
    ops.profile_enable(False)                       # reset launch counters
    step_resident(1)
    torch.cuda.synchronize()
    launches_per_step = sum(v["launches"] for v in ops.profile_read().values())
    if args.graph:                                  # capture outside the warm-up / timed regions (both input sets, both e2e slots)
        try:
            for k in range(2):
                step_resident_graph(k)
            torch.cuda.synchronize()
        except Exception as e:                      # capture is an optimisation of the launch path, never a requirement
            print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); falling back to eager launches", file=sys.stderr)
            graphs.clear()
            args.graph = 0
            torch.cuda.synchronize()
    pcie = None if args.no_extras else pcie_probe(dev, world, dist)
    resident = step_resident_graph if args.graph else step_resident
    for k in range(max(args.warmup, NSLOT)):            # at least one e2e step per slot before the timed region: each slot captures its graph on first use
        if k < args.warmup:
            resident(k)
        step_e2e(k)
    e2e_drain()
    torch.cuda.synchronize()
    ms_total = timed(resident, args.steps)
    launches = launches_per_step * args.steps
    ms_e2e = timed(step_e2e, args.steps, e2e_drain)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline leg: the same steps with CUDA events around every kernel launch (separate pass, not the `value` timing)
    # (levels run serially here: with one stream per level the kernels of the two levels overlap and per-launch event times
    # would include each other)
    ops.profile_enable(True)
    prof_steps = min(args.steps, 5)
    barrier()
    for k in range(prof_steps):
        hot_path(dev_in[k & 1], concurrent=False)
    barrier()
    prof = ops.profile_read()
    ops.profile_enable(False)

    if rank == 0:
        peaks = load_peaks()
        total_clips = clips * n_gpus * args.steps
        value = total_clips / (ms_total / 1e3)
        e2e_value = total_clips / (ms_e2e / 1e3)
        in_bytes = sum(t.numel() * t.element_size() for t in host_in[0])
        out_bytes = sum(t.numel() * t.element_size() for t in host_out2[0])
        tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
        dom_name, dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
        kernels = {k: {"ms_per_step": round(v["ms"] / prof_steps, 4), "share": round(v["ms"] / tot_ms, 4),
                       "launches_per_step": v["timed"] // prof_steps,
                       "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 and v["flops"] > 0 else None,
                       "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None}
                   for k, v in prof.items() if v["timed"]}
        achieved = dom["flops"] / (dom["ms"] * 1e-3) / 1e12 if dom["flops"] > 0 else dom["bytes"] / (dom["ms"] * 1e-3) / 1e9
        tensor_bound = dom["flops"] > 0 and dom_name in ("gemm_bf16_kernel", "spatial_attn_kernel", "spatial_attn_v2_kernel", "traj_fused_kernel",
                                                          "traj_ts_kernel", "ffn_fused_kernel", "ffn_n256_kernel", "qkv_fused_kernel", "qkv_direct_kernel", "traj_pair_kernel", "qkv_pair_kernel", "ffn_n256_pair_kernel")
        # the dominant kernel is event-timed launch by launch in a short pass at full clocks -> the BURST figure is its peak; the whole
        # step runs for seconds under the power cap -> its fraction is quoted against both
        peak = peaks["tf_burst"] if tensor_bound else peaks["hbm"]
        traffic, traffic_file = None, None
        for name in ("r02_traffic.json", "r01_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as f:
                    traffic = json.load(f).get(dom_name, {}).get("bytes_per_launch")
                traffic_file = name
                break
            except Exception:
                pass
        step_tf = flops_per_clip() * clips / (ms_total / args.steps * 1e-3) / 1e12
        roofline = {"kernel": dom_name, "bound": "tensor" if tensor_bound else "hbm", "achieved": round(achieved, 2), "peak": peak,
                    "unit": "TFLOP/s" if tensor_bound else "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                    "traffic_note": f"DRAM read+write bytes per launch of this kernel from the committed ncu --set full capture (profiles/{traffic_file})",
                    "peak_source": f"MEASURED_PEAKS.json ({peaks['source']}); burst bf16 figure: the kernel is timed alone, launch by launch",
                    "frac_of_sustained_peak": round(achieved / peaks["tf_sustained"], 4) if tensor_bound else None,
                    "avg_launch_ms": round(dom["ms"] / max(dom["timed"], 1), 5),
                    # the other roof of the same kernel: its algorithmic bytes over the same launches against the measured
                    # copy bandwidth, and where its arithmetic intensity sits relative to the ridge of the two measured peaks
                    "hbm_view": ({"achieved_gbs": round(dom["bytes"] / (dom["ms"] * 1e-3) / 1e9, 1), "peak_gbs": peaks["hbm"],
                                  "frac": round(dom["bytes"] / (dom["ms"] * 1e-3) / 1e9 / peaks["hbm"], 4),
                                  "intensity_flop_per_byte": round(dom["flops"] / dom["bytes"], 1),
                                  "ridge_flop_per_byte": round(peaks["tf_burst"] * 1e12 / (peaks["hbm"] * 1e9), 1),
                                  "note": "intensity below the ridge: by the roofline model this kernel's floor is its HBM time, not its tensor time"
                                          if dom["flops"] / dom["bytes"] < peaks["tf_burst"] * 1e12 / (peaks["hbm"] * 1e9) else
                                          "intensity above the ridge: the tensor roof binds"}
                                 if tensor_bound and dom["bytes"] > 0 else None),
                    "step_tflops": round(step_tf, 2),
                    "step_frac_of_burst_peak": round(step_tf / peaks["tf_burst"], 4),
                    "step_frac_of_tensor_peak": round(step_tf / peaks["tf_sustained"], 4),
                    "note": "per-kernel times from a separate pass with CUDA events around every launch, pyramid levels run serially",
                    "kernels": kernels}
        line = {"metric": METRIC, "value": round(value, 2), "unit": "clips/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": workload_config(clips),
                "e2e": {"value": round(e2e_value, 2), "unit": "clips/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                        "ms_per_step": round(ms_e2e / args.steps, 4),
                        "how": "nn.Module API, pinned host fp32 in/out every step, H2D / kernels / D2H pipelined on 3 streams over 3 staging slots" + (", kernels replayed from a CUDA graph" if args.graph else "")},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
        if pcie is not None:
            moved = (in_bytes + out_bytes) * n_gpus * args.steps / (ms_e2e * 1e-3) / 1e9
            line["e2e"]["roofline"] = {"h2d_gbs": pcie["h2d_gbs"], "d2h_gbs": pcie["d2h_gbs"], "box_ceiling_gbs": pcie["bidir_gbs"],
                                       "achieved_gbs": round(moved, 1), "frac": round(moved / pcie["bidir_gbs"], 3), "how": pcie["how"],
                                       "host_dtype": args.host_dtype, "d2h": args.d2h, "cpu_affinity_rank0": affinity}
        if n_gpus == 1 and not args.no_extras:
            line["module"] = module_figure(dev)
            line["gpu_eager_baseline"] = gpu_eager_baseline(clips, dev)
        if n_gpus == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
