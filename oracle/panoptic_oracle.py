"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's mask-wise panoptic post-processing.

Follows `MaXTronWCDeepLab.panoptic_mask_inference`, Vk/maxtron_deeplab/maxtron_wc_model.py:439-553 (the cross-clip model carries an
identical copy, Vk/maxtron_deeplab/maxtron_cc_model.py:460-574).  Pinned against the reference itself: `oracle/make_golden.py` runs the
unmodified method (imported through `oracle/ref_loader.wc_model`) and stores inputs + outputs in `tests/golden/panoptic_*.npz`;
`tests/test_oracle_golden.py` replays them through this file.  Only tests/, `__graft_entry__.smoke()` and bench.py's CPU legs may import it.

All decisions are threshold / ordering tests on fp32 softmax scores, so two implementations agree exactly whenever no score sits within
rounding distance of a threshold and no two reorder scores tie; `margins()` measures those distances so a test can state the precondition.
"""
from __future__ import annotations

import numpy as np


def _softmax(x: np.ndarray, axis: int) -> np.ndarray:
    x = x.astype(np.float32)
    e = np.exp(x - x.max(axis=axis, keepdims=True), dtype=np.float32)
    return (e / e.sum(axis=axis, keepdims=True, dtype=np.float32)).astype(np.float32)


class Metadata:
    """The three fields of the detectron2 metadata object the method reads (:466-473, :548)."""

    def __init__(self, thing_dataset_id_to_contiguous_id: dict, stuff_dataset_id_to_contiguous_id: dict, label_divisor: int):
        self.thing_dataset_id_to_contiguous_id = dict(thing_dataset_id_to_contiguous_id)
        self.stuff_dataset_id_to_contiguous_id = dict(stuff_dataset_id_to_contiguous_id)
        self.label_divisor = int(label_divisor)

    def tables(self, num_classes: int):
        """(cat_id[label], is_thing[label]) for label in [0, num_classes): `id_cont_to_ids_dic` (:469-473) and `label in thing_ids` (:489)."""
        thing_ids = list(self.thing_dataset_id_to_contiguous_id.values())
        stuff_ids = list(self.stuff_dataset_id_to_contiguous_id.values())
        all_ids = sorted(thing_ids + stuff_ids)
        if num_classes > len(all_ids):
            raise KeyError(f"{num_classes} classes but the metadata names only {len(all_ids)}")   # the reference raises KeyError at :524/:533
        cat = np.asarray(all_ids[:num_classes], dtype=np.int32)
        thing = np.asarray([1 if c in thing_ids else 0 for c in range(num_classes)], dtype=np.int32)
        return cat, thing


def scores(mask_cls: np.ndarray, mask_pred: np.ndarray, pixel_thr: float, w_cls: float = 1.0, w_mask: float = 1.0):
    """:456-467.  mask_cls [N, C+1], mask_pred [N, ...] -> (cls_scores, cls_labels, binary [N, P], pixel_number, reorder_score)."""
    N = mask_pred.shape[0]
    pc = _softmax(mask_cls, -1)[:, :-1]
    cls_scores, cls_labels = pc.max(-1), pc.argmax(-1)
    ms = _softmax(mask_pred.reshape(N, -1), 0)
    binary = ms > np.float32(pixel_thr)
    pixel_number = binary.sum(1).astype(np.float32)
    mask_score = (ms * binary).sum(1, dtype=np.float32) / np.maximum(pixel_number, np.float32(1.0))
    reorder = (cls_scores ** np.float32(w_cls)) * (mask_score ** np.float32(w_mask))
    return cls_scores, cls_labels, binary, pixel_number, reorder.astype(np.float32), ms


def margins(mask_cls, mask_pred, meta: Metadata, pixel_thr, thing_thr, stuff_thr, w_cls=1.0, w_mask=1.0):
    """Smallest distances to a decision boundary: (pixel score vs pixel_thr, class score vs its threshold, gap between neighbouring
    reorder scores relative to their size, gap between the two best class scores of a slot)."""
    cls_scores, cls_labels, _, pixel_number, reorder, ms = scores(mask_cls, mask_pred, pixel_thr, w_cls, w_mask)
    _, thing = meta.tables(mask_cls.shape[1] - 1)
    thr = np.where(thing[cls_labels] == 1, thing_thr, stuff_thr)
    srt = np.sort(reorder[pixel_number > 0].astype(np.float64))[::-1]      # slots without pixels are never accepted: their order is irrelevant
    rel_gap = np.min((srt[:-1] - srt[1:]) / np.maximum(srt[:-1], 1e-30)) if len(srt) > 1 else 1.0
    pc = np.sort(_softmax(mask_cls, -1)[:, :-1], axis=-1)
    top2 = np.min(pc[:, -1] - pc[:, -2]) if pc.shape[1] > 1 else 1.0
    return float(np.abs(ms - pixel_thr).min()), float(np.abs(cls_scores - thr).min()), float(rel_gap), float(top2)


def panoptic_mask_inference(mask_cls, mask_pred, mask_embedding, meta: Metadata, pixel_thr=0.3, thing_thr=0.1, stuff_thr=0.3,
                            overlap_thr=0.8, w_cls=1.0, w_mask=1.0):
    """Returns (panoptic_seg_mask int32 [T,H,W], dic_cat_idemb {cat_id: [L2-normalised embedding, ...]}, segments), where `segments`
    lists (slot, label, is_thing, final_id) of every slot that opened a segment, in acceptance order (what `segments_info` and
    `dic_tmp` record, :516-540)."""
    mask_cls = np.asarray(mask_cls, dtype=np.float32)
    mask_pred = np.asarray(mask_pred, dtype=np.float32)
    N = mask_pred.shape[0]
    out_shape = mask_pred.shape[1:]
    cls_scores, cls_labels, binary, pixel_number, reorder, _ = scores(mask_cls, mask_pred, pixel_thr, w_cls, w_mask)
    order = np.argsort(-reorder, kind="stable")                                   # :468 (descending)
    cat, thing = meta.tables(mask_cls.shape[1] - 1)
    seg = np.full(binary.shape[1], -1, dtype=np.int32)                            # -1 = the reference's `panoptic_seg == 0` (unassigned)
    per_cat_count: dict = {}
    stuff_seen = set()
    segments = []
    for i in range(N):                                                            # :486-540
        cur = int(order[i])
        label = int(cls_labels[cur])
        is_thing = bool(thing[label])
        confident = cls_scores[cur] > np.float32(thing_thr if is_thing else stuff_thr)
        new_mask = binary[cur] & (seg == -1)
        ok = np.float32(new_mask.sum()) > np.float32(pixel_number[cur] * np.float32(overlap_thr))
        if not (confident and ok):
            continue
        cid = int(cat[label])
        if is_thing:
            ii = per_cat_count.get(cid, 0)
            per_cat_count[cid] = ii + 1
            final = cid * meta.label_divisor + ii                                 # :548
            segments.append((cur, label, 1, final))
        else:
            final = cid                                                           # :553; merged stuff regions share the id (:506-511)
            if label not in stuff_seen:
                stuff_seen.add(label)
                segments.append((cur, label, 0, final))
        seg[new_mask] = final
    dic = {}
    if mask_embedding is not None:
        emb = np.asarray(mask_embedding, dtype=np.float32)
        for slot, label, is_thing, final in segments:
            if is_thing:
                v = emb[slot]
                dic.setdefault(int(cat[label]), []).append(v / max(float(np.sqrt((v * v).sum(dtype=np.float32))), 1e-12))   # F.normalize (:549)
    return seg.reshape(out_shape), dic, segments
