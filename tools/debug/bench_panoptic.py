"""Debug: time the GPU-resident panoptic merge (row f4) at the config size (128 slots, 124 classes, 2 x 641 x 641) against the reference's
own formulation (a per-slot loop of full-frame torch ops with `.item()` syncs) run on the same GPU.
usage: python tools/debug/bench_panoptic.py [H W]"""
import sys, time, types, torch
sys.path.insert(0, ".")
from axial_vs_b200 import synth
from axial_vs_b200.postprocess import PanopticPostProcessor

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (641, 641)
N, C, T = 128, 124, 2
thing, stuff, div = synth.panoptic_metadata(C)
md = types.SimpleNamespace(thing_dataset_id_to_contiguous_id=thing, stuff_dataset_id_to_contiguous_id=stuff, label_divisor=div)
pp = PanopticPostProcessor(md)
mc, mp, me = (t.cuda() for t in synth.panoptic_case(7, N, C, T, H, W, cell=H // 12))


def loop_formulation(mask_cls, mask_pred):
    """The reference's per-slot loop (maxtron_wc_model.py:456-540) with torch ops on the GPU, final ids painted directly."""
    F = torch.nn.functional
    cls_scores, cls_labels = F.softmax(mask_cls, dim=-1)[..., :-1].max(-1)
    ms = F.softmax(mask_pred, dim=0)
    binary = ms > pp.pixel_confidence_threshold
    bf = binary.flatten(1).float()
    pn = bf.sum(1)
    score = (cls_scores ** 1.0) * ((ms.flatten(1) * bf).sum(1) / torch.clamp(pn, min=1.0))
    order = torch.argsort(score, descending=True)
    seg = torch.zeros(mask_pred.shape[1:], dtype=torch.int32, device=mask_pred.device)
    thing_ids = set(thing.values())
    cur_id, memory = 0, {}
    for i in range(N):
        cur = order[i].item()
        s, lab = cls_scores[cur].item(), cls_labels[cur].item()
        is_thing = lab in thing_ids
        conf = s > (0.1 if is_thing else 0.3)
        orig = binary[cur].float().sum()
        new = torch.logical_and(binary[cur], seg == 0)
        ok = new.float().sum() > orig * 0.8
        if conf and ok:
            if not is_thing:
                if lab in memory:
                    seg[new] = memory[lab]
                    continue
                memory[lab] = cur_id + 1
            cur_id += 1
            seg[new] = cur_id
    return seg


for _ in range(3):
    pp.panoptic_segments(mc, mp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    seg, segs = pp.panoptic_segments(mc, mp)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
px = T * H * W
print(f"panoptic merge, {N} slots x {T}x{H}x{W}: {ms * 1e3:.1f} us  ({N * px * 4 / ms / 1e6:.0f} GB/s of logits), {int(segs[0])} segments, "
      f"{(seg >= 0).float().mean().item() * 100:.1f} % of the pixels assigned")
t0 = time.time(); pp.panoptic_mask_inference(mc, mp, me); torch.cuda.synchronize()
print(f"  drop-in call incl. the one host read and the dict: {(time.time() - t0) * 1e3:.2f} ms")
loop_formulation(mc, mp); torch.cuda.synchronize()
t0 = time.time(); loop_formulation(mc, mp); torch.cuda.synchronize()
print(f"  per-slot loop formulation (torch ops on this GPU, 3 host syncs per slot): {(time.time() - t0) * 1e3:.2f} ms")
from axial_vs_b200 import _lib
lib = _lib.load()
