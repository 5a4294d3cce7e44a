"""bench.py --config {3,4,5}: the BASELINE.json configurations besides the metric's own (configs[1], bench.py itself).

Same JSON line as bench.py (metric, value, e2e, roofline of the dominant kernel, clocks, gpu_launches); each configuration names its
workload in `config.workload`.  Inputs are synthetic (N(0,1) features, xavier random-init weights); steps are launched eagerly.

  3  cross-clip tracking module (CC:204-331) on ONE video of 64 clips x 2 frames, Q = 128 queries, 4 layers, 641x641 frames -> 161x161
     mask features.  N GPUs share the video clip-wise (`CrossClipTrackingModule.forward_sharded`: one NCCL all-gather of the clip queries,
     the layers run redundantly, mask logits per rank, one all-gather of the logits) -> STRONG scaling; value = clips of the video / s.
  4  Tube-Link flavour (TL/mmdet/models/plugins/msdeformattn_pixel_decoder.py:616-632): T = 5 frames of 480x640 -> temporal levels
     15x20 and 30x40, 6 encoder layers x [f + gamma * TemporalEncoder(1 axial layer)(f, pos3d)], gamma = 1; weak scaling over clips.
  5  axial-trajectory micro-benchmark sweep: T in {2,5,10} x H=W in {41,81,161}, one axial layer per point, one step = all nine
     points once (batch per point chosen for ~4 waves of 128-row tiles); weak scaling; per-point figures in `sweep`.
"""
from __future__ import annotations

import json
import os
import sys
import time

import torch

import bench as B


def _env():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: a CUDA device is required (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return dist, world, rank, local_rank, dev


def _timed(fn, steps, dist, world, dev, drain=None):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        fn(k)
    if drain is not None:
        drain()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


def _profile(fn, steps, ops):
    ops.profile_enable(True)
    for k in range(steps):
        fn(k)
    torch.cuda.synchronize()
    prof = ops.profile_read()
    ops.profile_enable(False)
    return prof


def _roofline(prof, prof_steps, peaks, step_flops, ms_per_step):
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    timed = {k: v for k, v in prof.items() if v["timed"]}
    dom_name, dom = max(timed.items(), key=lambda kv: kv[1]["ms"])
    kernels = {k: {"ms_per_step": round(v["ms"] / prof_steps, 4), "share": round(v["ms"] / tot_ms, 4), "launches_per_step": v["timed"] // prof_steps,
                   "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 and v["flops"] > 0 else None,
                   "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None} for k, v in timed.items()}
    tensor_bound = dom["flops"] > 0 and dom_name in ("gemm_bf16_kernel", "traj_ts_kernel", "traj_fused_kernel", "ffn_fused_kernel", "ffn_n256_kernel",
                                                      "qkv_fused_kernel", "qkv_direct_kernel", "traj_pair_kernel", "qkv_pair_kernel", "ffn_n256_pair_kernel")
    achieved = dom["flops"] / (dom["ms"] * 1e-3) / 1e12 if tensor_bound else dom["bytes"] / (dom["ms"] * 1e-3) / 1e9
    peak = peaks["tf_burst"] if tensor_bound else peaks["hbm"]
    step_tf = step_flops / (ms_per_step * 1e-3) / 1e12
    return {"kernel": dom_name, "bound": "tensor" if tensor_bound else "hbm", "achieved": round(achieved, 2), "peak": peak,
            "unit": "TFLOP/s" if tensor_bound else "GB/s", "frac": round(achieved / peak, 4), "traffic": None,
            "peak_source": f"MEASURED_PEAKS.json ({peaks['source']}); burst bf16 figure / copy bandwidth: the kernel is timed alone, launch by launch",
            "avg_launch_ms": round(dom["ms"] / max(dom["timed"], 1), 5), "step_tflops": round(step_tf, 2),
            "step_frac_of_burst_peak": round(step_tf / peaks["tf_burst"], 4), "step_frac_of_tensor_peak": round(step_tf / peaks["tf_sustained"], 4),
            "note": "per-kernel times from a separate pass with CUDA events around every launch", "kernels": kernels}


def _emit(args, rank, world, sampler, metric_unit, value, ms_total, e2e, launches, roofline, config, scaling, extra=None):
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return
    line = {"metric": B.METRIC, "value": round(value, 2), "unit": metric_unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": config, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def _sampler(rank, local_rank):
    s = B.ClockSampler(local_rank)
    if rank == 0:
        s.start()
        time.sleep(0.6)
    return s


# ------------------------------------------------------------------------------------------------ config 3: cross-clip module, 64 clips
def flops_cc_layer(T, Q, C=256):
    N = T * Q
    ta = N * C * (10 * C + 4 * T * C + 4 * N + 4 * T)                 # SURVEY.md section 8d, B' = 1, F = T clips
    aspp = N * (3 * 2 * 768 * C + 2 * 768 * C)                        # three k=3 convs + the 768 -> 256 projection
    return ta + aspp


def run_cfg3(args):
    from axial_vs_b200 import cross_clip, ops, sharding, synth
    dist, world, rank, local_rank, dev = _env()
    Q, T, V, L, K = 128, 64, 2, 4, 124
    Hm = Wm = 161                                                      # 641 -> 161 (stride 4) mask features
    if T % world:
        raise SystemExit(f"--config 3: {T} clips are split evenly over the ranks (got {world})")
    t0, t1 = sharding.shard_range(T, rank, world)
    Tl = t1 - t0
    m = cross_clip.CrossClipTrackingModule(num_layers=L, num_classes=K, attn_drop=0.0, aspp_drop=0.0, kernel_sizes=[3, 3, 3], atrous_rates=[1, 2, 3],
                                           norm_fn="ln", num_clip_frames=V).eval()
    m.load_state_dict(synth.cross_clip_params(11, L, K), strict=True)
    m.to(dev)
    # within-clip stage of the same video: the temporal hot path (bench.py's workload) over THIS rank's clips
    from axial_vs_b200 import modules, within_clip
    encoders = []
    for s_ in range(B.STAGES):
        enc = modules.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "axial-trajectory", B.LAYERS_PER_STAGE).eval()
        enc.load_state_dict(synth.encoder_params(s_, B.LAYERS_PER_STAGE), strict=True)
        encoders.append(enc.to(dev))
    le = synth.level_embed(99).to(dev)
    wc_pos = [ops.pos3d(1, V, H, W, le[i].contiguous(), dev).expand(Tl, -1, -1, -1, -1) for i, (H, W) in enumerate(B.LEVELS)]
    wc_in = [torch.randn(Tl * V, H * W, 256, device=dev) for (H, W) in B.LEVELS]

    @torch.no_grad()
    def within_clip_stage():
        cur = list(wc_in)
        for enc in encoders:
            cur = [o[0] for o in within_clip.run_levels_concurrent(enc, cur, wc_pos, 1)]
        return cur

    g = torch.Generator().manual_seed(100 + rank)
    h_cq = torch.randn(1, Q, Tl, 256, generator=g).pin_memory()                     # this rank's clips (cluster centres of the clip segmenter)
    h_pf = torch.randn(1, 128, Tl * V, Hm, Wm, generator=g).pin_memory()            # and their pixel features
    d_cq, d_pf = h_cq.to(dev), h_pf.to(dev)
    h_out = torch.empty(Q, Tl * V, Hm, Wm).pin_memory()

    @torch.no_grad()
    def step(k, cq=None, pf=None):
        cq, pf = (d_cq, d_pf) if cq is None else (cq, pf)
        within_clip_stage()                                            # clips are independent: no collective
        return m.forward_sharded(cq, pf, T)                            # final-layer predictions only (N = 1: no collective is issued)

    @torch.no_grad()
    def step_e2e(k):
        out = step(k, h_cq.to(dev, non_blocking=True), h_pf.to(dev, non_blocking=True))
        h_out.copy_(out["pred_masks"][0, :, t0 * V:t1 * V], non_blocking=True)       # the reference's `.cpu()` of the mask logits (CC:70), own clips
        return out

    sampler = _sampler(rank, local_rank)
    step(0)
    torch.cuda.synchronize()
    ops.profile_enable(False)
    step(1)
    torch.cuda.synchronize()
    launches = sum(v["launches"] for v in ops.profile_read().values())
    for k in range(max(args.warmup, 3)):
        step(k)
        step_e2e(k)
    ms_total = _timed(step, args.steps, dist, world, dev)
    ms_e2e = _timed(step_e2e, args.steps, dist, world, dev)
    prof_steps = min(args.steps, 3)
    prof = _profile(step, prof_steps, ops)
    step_flops = L * flops_cc_layer(T, Q) + 2.0 * T * Q * (V * Hm * Wm) * 128     # layers (run redundantly per rank, counted once) + the mask contraction
    step_flops += B.flops_per_clip() * T                                          # + the within-clip hot path of the 64 clips
    peaks = B.load_peaks()
    roof = _roofline(prof, prof_steps, peaks, step_flops, ms_total / args.steps) if rank == 0 else None
    e2e = {"value": round(T * args.steps / (ms_e2e * 1e-3), 2), "unit": "clips/s", "h2d_bytes_per_step": h_cq.numel() * 4 + h_pf.numel() * 4,
           "d2h_bytes_per_step": h_out.numel() * 4, "ms_per_step": round(ms_e2e / args.steps, 4),
           "how": "CrossClipTrackingModule API; per step every rank copies its clips' queries and pixel features from pinned host memory and reads "
                  "its clips' mask logits back (bytes are per rank)"}
    cfg = {"workload": "within-clip + cross-clip tracking on one long synthetic video (BASELINE configs[2]): 64 clips x 2 frames clip-sharded over the "
                       "GPUs; per step every rank runs the within-clip temporal hot path on its clips (res5 21x21 + res4 41x41, no collective), then "
                       "the cross-clip module (Q = 128, 4 layers, mask features 128 x 161 x 161 per frame) through forward_sharded: all-gather of the "
                       "clip queries, layers run redundantly, per-rank mask logits, all-gather of the logits",
           "clips_per_video": T, "clips_per_gpu": Tl, "gflop_per_step": round(step_flops / 1e9, 1),
           "l2": "per-step pixel features and mask logits (1.7 GB each) exceed the 126 MB L2"}
    _emit(args, rank, world, sampler, "clips/s", T * args.steps / (ms_total * 1e-3), ms_total, e2e, launches * args.steps, roof, cfg, "strong")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ config 4: Tube-Link T=5
def flops_axial_layer(Bc, T, H, W, C=256, dffn=1024):
    return Bc * T * H * W * C * (20 * C + 8 * T * C + 4 * T * (H + W) + 8 * T + 4 * dffn)


def run_cfg4(args):
    from axial_vs_b200 import ops, synth, tube_link
    dist, world, rank, local_rank, dev = _env()
    T, levels, n_layers = 5, [(15, 20), (30, 40)], 6
    clips = args.clips if args.clips != 42 else 16
    layers = []
    for i in range(n_layers):
        enc = tube_link.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, 1).eval()
        enc.load_state_dict(synth.encoder_params(40 + i, 1), strict=True)
        layers.append(enc.to(dev))
    gamma = torch.ones(256, device=dev)                                # the 1e-6 init makes the branch numerically invisible
    le = synth.level_embed(77).to(dev)
    pos = [ops.pos3d(1, T, H, W, le[i].contiguous(), dev).expand(clips, -1, -1, -1, -1) for i, (H, W) in enumerate(levels)]
    h_in = [torch.randn(clips * T, H * W, 256, generator=torch.Generator().manual_seed(300 + rank + i)).pin_memory() for i, (H, W) in enumerate(levels)]
    d_in = [t.to(dev) for t in h_in]
    h_out = [torch.empty_like(t).pin_memory() for t in h_in]

    @torch.no_grad()
    def step(k, feats=None):
        cur = list(d_in if feats is None else feats)
        for enc in layers:                                             # one encoder layer = MSDA sampling (not on this path) + the temporal branch
            cur = tube_link.temporal_branch(cur, pos, enc, gamma, len(levels))
        return cur

    @torch.no_grad()
    def step_e2e(k):
        outs = step(k, [t.to(dev, non_blocking=True) for t in h_in])
        for o, h in zip(outs, h_out):
            h.copy_(o, non_blocking=True)
        return outs

    sampler = _sampler(rank, local_rank)
    step(0)
    torch.cuda.synchronize()
    ops.profile_enable(False)
    step(1)
    torch.cuda.synchronize()
    launches = sum(v["launches"] for v in ops.profile_read().values())
    for k in range(max(args.warmup, 3)):
        step(k)
        step_e2e(k)
    ms_total = _timed(step, args.steps, dist, world, dev)
    ms_e2e = _timed(step_e2e, args.steps, dist, world, dev)
    prof_steps = min(args.steps, 3)
    prof = _profile(step, prof_steps, ops)
    step_flops = n_layers * sum(flops_axial_layer(clips, T, H, W) for H, W in levels)
    roof = _roofline(prof, prof_steps, B.load_peaks(), step_flops, ms_total / args.steps) if rank == 0 else None
    nb = sum(t.numel() * 4 for t in h_in)
    e2e = {"value": round(clips * world * args.steps / (ms_e2e * 1e-3), 2), "unit": "clips/s", "h2d_bytes_per_step": nb, "d2h_bytes_per_step": nb,
           "ms_per_step": round(ms_e2e / args.steps, 4), "how": "tube_link module API, pinned host fp32 in / out every step (bytes per rank)"}
    cfg = {"workload": "Tube-Link + MaXTron temporal branch (BASELINE configs[3]): T = 5, 480x640 -> levels 15x20 + 30x40, 6 encoder layers x "
                       "[f + gamma * TemporalEncoder(1 axial-trajectory layer)(f, pos3d)], gamma = 1",
           "clips_per_gpu_per_step": clips, "gflop_per_clip": round(step_flops / clips / 1e9, 2),
           "l2": f"{sum(t.numel() for t in d_in) * 4 / 1e6:.0f} MB of activations per step"}
    _emit(args, rank, world, sampler, "clips/s", clips * world * args.steps / (ms_total * 1e-3), ms_total, e2e, launches * args.steps, roof, cfg, "weak")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ config 5: the sweep
def run_cfg5(args):
    from axial_vs_b200 import ops, synth
    from axial_vs_b200.modules import TemporalAxialTrajectoryAttentionLayer
    dist, world, rank, local_rank, dev = _env()
    layer = TemporalAxialTrajectoryAttentionLayer(256, 1024, 0.0, 0.0, "relu", 8).eval()
    layer.load_state_dict(synth.axial_layer_params(0))
    layer.to(dev)
    points = []
    for T in (2, 5, 10):
        for HW in (41, 81, 161):
            tokens = T * HW * HW
            Bc = max(1, min(64, (4 * 148 * 128 + tokens - 1) // tokens))
            src = torch.randn(Bc * T, HW * HW, 256, device=dev)
            pos = torch.randn(1, T, HW, HW, 256, device=dev).expand(Bc, -1, -1, -1, -1)
            points.append((T, HW, Bc, src, pos))

    @torch.no_grad()
    def step(k):
        return [layer(src, pos)[0] for (_, _, _, src, pos) in points]

    sampler = _sampler(rank, local_rank)
    step(0)
    torch.cuda.synchronize()
    ops.profile_enable(False)
    step(1)
    torch.cuda.synchronize()
    launches = sum(v["launches"] for v in ops.profile_read().values())
    for k in range(max(args.warmup, 3)):
        step(k)
    ms_total = _timed(step, args.steps, dist, world, dev)
    # per point (rank 0's device, eager, events around the layer)
    sweep = []
    peaks = B.load_peaks()
    for (T, HW, Bc, src, pos) in points:
        with torch.no_grad():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                layer(src, pos)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        fl = flops_axial_layer(Bc, T, HW, HW)
        sweep.append({"T": T, "HW": HW, "clips": Bc, "ms_per_layer": round(ms, 3), "tflops": round(fl / (ms * 1e-3) / 1e12, 1),
                      "frac_of_burst_peak": round(fl / (ms * 1e-3) / 1e12 / peaks["tf_burst"], 3)})
    # e2e of the sweep: the smallest and the largest point through host buffers would only measure PCIe; report the aggregate through host
    # memory for the T=2, 41x41 point (the metric's shape)
    T0, HW0, B0, src0, pos0 = points[0]
    h_in, h_out = src0.cpu().pin_memory(), torch.empty_like(src0, device="cpu").pin_memory()

    @torch.no_grad()
    def step_e2e(k):
        o = layer(h_in.to(dev, non_blocking=True), pos0)[0]
        h_out.copy_(o, non_blocking=True)

    for k in range(3):
        step_e2e(k)
    ms_e2e = _timed(step_e2e, args.steps, dist, world, dev)
    prof_steps = min(args.steps, 3)
    prof = _profile(step, prof_steps, ops)
    step_flops = sum(flops_axial_layer(Bc, T, HW, HW) for (T, HW, Bc, _, _) in points)
    roof = _roofline(prof, prof_steps, peaks, step_flops, ms_total / args.steps) if rank == 0 else None
    clips_per_step = sum(p[2] for p in points)
    e2e = {"value": round(B0 * world * args.steps / (ms_e2e * 1e-3), 2), "unit": "clips/s", "h2d_bytes_per_step": h_in.numel() * 4,
           "d2h_bytes_per_step": h_out.numel() * 4, "ms_per_step": round(ms_e2e / args.steps, 4),
           "how": f"one axial layer at the T=2, 41x41 point ({B0} clips) through pinned host fp32 in / out; the other points are device-timed only"}
    cfg = {"workload": "axial-trajectory attention micro-benchmark sweep (BASELINE configs[4]): T in {2,5,10} x H=W in {41,81,161}, C=256, 8 heads, one "
                       "TemporalAxialTrajectoryAttentionLayer per point; one step = the nine points once",
           "clips_per_gpu_per_step": clips_per_step, "gflop_per_step": round(step_flops / 1e9, 1), "l2": "activations of every point exceed the L2 except T=2, 41x41"}
    _emit(args, rank, world, sampler, "clips/s", clips_per_step * world * args.steps / (ms_total * 1e-3), ms_total, e2e, launches * args.steps, roof, cfg,
          "weak", {"sweep": sweep})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run(args):
    {3: run_cfg3, 4: run_cfg4, 5: run_cfg5}[args.config](args)
