#!/usr/bin/env python
"""bench.py -- clips/s of the within-clip tracking module's temporal hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clips C] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Video-kMaX R50 + MaXTron within-clip tracking on VIPSeg-shaped synthetic clips,
T=2 frames of 641x641 -> temporal levels res5 (21x21) and res4 (41x41), C=256, 2 stages x TemporalEncoder(2 axial-
trajectory layers), random-init weights (xavier), synthetic N(0,1) features.  One step = that hot path over one
batch of `--clips` clips per GPU (weak scaling: every rank processes its own clips, no collective in the data path;
a small per-clip output summary is all-gathered once per step).  Algorithmic work: 60.5 GFLOP per clip.

`--impl reference` times the reference's CPU implementation of the same path (the oracle port checked against the
reference in tests/; the Python reference tree itself cannot travel to the GPU box) on the host cores.
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
def oracle_clip_runner():
    """Returns fn() that runs the hot path of ONE clip on the CPU (fp32 oracle port) and the thread count used."""
    from axial_vs_b200 import synth
    from oracle import traj_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    stages = [O.split_encoder_params(synth.encoder_params(s, LAYERS_PER_STAGE)) for s in range(STAGES)]
    le = synth.level_embed(99)
    srcs = [synth.randn(10 + i, T_FRAMES, H * W, 256) for i, (H, W) in enumerate(LEVELS)]
    poss = [O.level_pos3d(1, T_FRAMES, H, W, le[i]) for i, (H, W) in enumerate(LEVELS)]

    @torch.no_grad()
    def run():
        cur = list(srcs)
        for st in stages:
            for i in range(len(LEVELS)):
                cur[i], _, _ = O.temporal_encoder(cur[i], poss[i], st)
        return cur

    return run, torch.get_num_threads()


def cpu_baseline_sample(budget_s: float = 12.0):
    run, cores = oracle_clip_runner()
    run()                                  # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        run()
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 200:
            break
    return {"value": n / el, "unit": "clips/s", "cores": cores, "kind": "port",
            "sample": f"{n} clips (T=2, res5 21x21 + res4 41x41, 2 stages x 2 layers) in {el:.1f} s, fp32 torch-CPU oracle port"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run, cores = oracle_clip_runner()
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
            "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": "port",
                             "sample": "bounded sample of the workload: 1 clip per timed step (the CPU port is batch-size independent: clips run "
                                       "one after the other); the oracle port of the reference modules (oracle/traj_oracle.py), all host threads"},
            "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(clips):
    return {"workload": "Video-kMaX R50 + MaXTron within-clip tracking hot path (BASELINE configs[1]): T=2, 641x641 -> "
                        "res5 21x21 + res4 41x41, C=256, 2 stages x TemporalEncoder(2 axial-trajectory layers)",
            "clips_per_gpu_per_step": clips, "gflop_per_clip": round(flops_per_clip() / 1e9, 2),
            "pos": "one PositionEmbeddingSine3D table per level shared by all clips (the reference's table is clip-independent)",
            "l2": "per-step activations exceed the 126 MB L2 and two input sets alternate"}


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
    ap.add_argument("--clip-chunks", type=int, default=1, help="split each level's clips into this many groups, one CUDA stream per (level, group)")
    ap.add_argument("--full-pos", type=int, default=0, help="1: materialise the positional table for every clip ([B,T,H,W,C], as the reference passes it) instead of sharing one table")
    ap.add_argument("--level-streams", type=int, default=1, help="1: run the two pyramid levels on two CUDA streams (default), 0: serially")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
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
    host_in = [[torch.randn(clips * T_FRAMES, H * W, 256, generator=torch.Generator().manual_seed(1000 * k + 10 * rank + i)).pin_memory()
                for i, (H, W) in enumerate(LEVELS)] for k in range(2)]
    dev_in = [[t.to(dev) for t in hs] for hs in host_in]
    host_out = host_in[0]   # shapes of the per-step outputs (for byte accounting)

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
    # loop would: H2D of step k+1 and D2H of step k-1 overlap the kernels of step k (double-buffered pinned host + device
    # staging); every step still moves its own inputs in and its own outputs out inside the timed region.
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    stage_in = [[torch.empty_like(t, device=dev) for t in host_in[0]] for _ in range(2)]
    host_out2 = [[torch.empty_like(t).pin_memory() for t in host_in[0]] for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]        # H2D of slot finished
    ev_used = [torch.cuda.Event() for _ in range(2)]      # compute finished reading slot's staging inputs
    ev_comp = [torch.cuda.Event() for _ in range(2)]      # compute of slot finished
    ev_out = [torch.cuda.Event() for _ in range(2)]       # D2H of slot finished
    e2e_state = {"n": 0, "keep": [None, None]}
    e2e_graphs = {}

    def step_e2e(k):
        slot = e2e_state["n"] & 1
        first = e2e_state["n"] < 2
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
            # rewritten two steps later, after the D2H of this slot has been waited for
            if not first:
                cur.wait_event(ev_out[slot])
            g = e2e_graphs.get(slot)
            if g is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    e2e_graphs[("out", slot)] = hot_path(stage_in[slot])
                e2e_graphs[slot] = g
            g.replay()
            outs = e2e_graphs[("out", slot)]
        else:
            outs = hot_path(stage_in[slot])
        ev_used[slot].record(cur)
        gather_summary(outs)
        ev_comp[slot].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_comp[slot])
            if not first:
                pass                                   # host_out2[slot] was drained two steps ago (same stream, in order)
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
        e0.record()
        for k in range(steps):
            fn(k)
        if drain is not None:
            drain()                                    # the timed region ends when the last D2H has landed
        e1.record()
        barrier()
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
    resident = step_resident_graph if args.graph else step_resident
    for k in range(args.warmup):
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
        in_bytes = sum(t.numel() * 4 for t in host_in[0])
        out_bytes = sum(t.numel() * 4 for t in host_out)
        tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
        dom_name, dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
        kernels = {k: {"ms_per_step": round(v["ms"] / prof_steps, 4), "share": round(v["ms"] / tot_ms, 4),
                       "launches_per_step": v["timed"] // prof_steps,
                       "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 and v["flops"] > 0 else None,
                       "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None}
                   for k, v in prof.items() if v["timed"]}
        achieved = dom["flops"] / (dom["ms"] * 1e-3) / 1e12 if dom["flops"] > 0 else dom["bytes"] / (dom["ms"] * 1e-3) / 1e9
        tensor_bound = dom["flops"] > 0 and dom_name in ("gemm_bf16_kernel", "spatial_attn_kernel", "spatial_attn_v2_kernel", "traj_fused_kernel",
                                                          "traj_ts_kernel", "ffn_fused_kernel", "ffn_n256_kernel", "qkv_fused_kernel")
        peak = peaks["tf_sustained"] if tensor_bound else peaks["hbm"]
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
                traffic = json.load(f).get(dom_name, {}).get("bytes_per_launch")
        except Exception:
            pass
        roofline = {"kernel": dom_name, "bound": "tensor" if tensor_bound else "hbm", "achieved": round(achieved, 2), "peak": peak,
                    "unit": "TFLOP/s" if tensor_bound else "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                    "traffic_note": "DRAM read+write bytes per launch of this kernel from the committed ncu --set full capture (profiles/r01_traffic.json)",
                    "peak_source": f"MEASURED_PEAKS.json ({peaks['source']}; sustained bf16 figure: kernel timed inside a long step)",
                    "avg_launch_ms": round(dom["ms"] / max(dom["timed"], 1), 5),
                    "step_tflops": round(flops_per_clip() * clips / (ms_total / args.steps * 1e-3) / 1e12, 2),
                    "step_frac_of_tensor_peak": round(flops_per_clip() * clips / (ms_total / args.steps * 1e-3) / 1e12 / peaks["tf_sustained"], 4),
                    "note": "per-kernel times from a separate pass with CUDA events around every launch, pyramid levels run serially",
                    "kernels": kernels}
        line = {"metric": METRIC, "value": round(value, 2), "unit": "clips/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": workload_config(clips),
                "e2e": {"value": round(e2e_value, 2), "unit": "clips/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                        "ms_per_step": round(ms_e2e / args.steps, 4),
                        "how": "nn.Module API, pinned host fp32 in/out every step, H2D / kernels / D2H pipelined on 3 streams" + (", kernels replayed from a CUDA graph" if args.graph else "")},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
        if n_gpus == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
