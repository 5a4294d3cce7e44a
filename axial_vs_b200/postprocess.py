"""Post-path tail: the mask-wise panoptic merge, GPU-resident.

Drop-in for `MaXTronWCDeepLab.panoptic_mask_inference` / `MaXTronCCDeepLab.panoptic_mask_inference`
(Vk/maxtron_deeplab/maxtron_wc_model.py:439-553, maxtron_cc_model.py:460-574): same arguments, same return value
(`panoptic_seg_mask` int32 [T, H, W], `dic_cat_idemb` {category id: [L2-normalised mask embedding, ...]}).  The reference walks the 128
mask slots in a Python loop with three `.item()` syncs and several full-frame kernels per slot; here five kernels run back to back and
the host reads one small segment table at the end (needed to build the returned dict).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Tuple

import torch

from . import _lib, ops


class PanopticPostProcessor:
    """Holds what the reference method reads from `self` (:443-448, :466-473, :548): the six thresholds / weights and the metadata's
    `thing_dataset_id_to_contiguous_id`, `stuff_dataset_id_to_contiguous_id`, `label_divisor`."""

    def __init__(self, metadata, pixel_confidence_threshold: float = 0.3, class_threshold_thing: float = 0.1, class_threshold_stuff: float = 0.3,
                 overlap_threshold: float = 0.8, reorder_class_weight: float = 1.0, reorder_mask_weight: float = 1.0):
        self.metadata = metadata
        self.pixel_confidence_threshold = float(pixel_confidence_threshold)
        self.class_threshold_thing = float(class_threshold_thing)
        self.class_threshold_stuff = float(class_threshold_stuff)
        self.overlap_threshold = float(overlap_threshold)
        self.reorder_class_weight = float(reorder_class_weight)
        self.reorder_mask_weight = float(reorder_mask_weight)
        self._tables: Dict[Tuple[int, str], Tuple[torch.Tensor, torch.Tensor]] = {}

    def _label_tables(self, num_classes: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        key = (num_classes, str(device))
        if key not in self._tables:
            thing_ids = list(self.metadata.thing_dataset_id_to_contiguous_id.values())
            stuff_ids = list(self.metadata.stuff_dataset_id_to_contiguous_id.values())
            all_ids = sorted(thing_ids + stuff_ids)                       # id_cont_to_ids_dic[ii] = all_ids[ii]  (:469-473)
            if num_classes > len(all_ids):
                raise KeyError(f"mask_cls has {num_classes} classes but the metadata names only {len(all_ids)}")
            cat = torch.tensor(all_ids[:num_classes], dtype=torch.int32, device=device)
            thing = torch.tensor([1 if c in thing_ids else 0 for c in range(num_classes)], dtype=torch.int32, device=device)
            self._tables[key] = (cat, thing)
        return self._tables[key]

    @torch.no_grad()
    def panoptic_segments(self, mask_cls: torch.Tensor, mask_pred: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(panoptic_seg_mask int32 [T,H,W], segments int32 [1 + 4N]) -- both stay on the device, no host sync."""
        if mask_pred.device.type != "cuda":
            raise RuntimeError("axial_vs_b200: CUDA tensors required (there is no CPU fallback)")
        if mask_cls.dim() != 2 or mask_pred.dim() < 2 or mask_cls.shape[0] != mask_pred.shape[0]:
            raise RuntimeError(f"mask_cls must be [N, C+1] and mask_pred [N, ...] (got {tuple(mask_cls.shape)}, {tuple(mask_pred.shape)})")
        N, C1 = mask_cls.shape
        out_shape = tuple(mask_pred.shape[1:])
        mc = mask_cls.contiguous().float()
        mp = mask_pred.contiguous().float()
        P = mp.numel() // N
        cat, thing = self._label_tables(C1 - 1, mp.device)
        lib = _lib.load()
        seg = torch.empty(P, dtype=torch.int32, device=mp.device)
        segments = torch.zeros(1 + 4 * N, dtype=torch.int32, device=mp.device)
        nbytes = lib.axvs_panoptic_workspace_bytes(N, P)
        with torch.cuda.device(mp.device):
            ws = ops.workspace(nbytes, mp.device)
            rc = lib.axvs_panoptic_inference(mc.data_ptr(), mp.data_ptr(), N, C1 - 1, P, cat.data_ptr(), thing.data_ptr(), int(self.metadata.label_divisor),
                                             self.pixel_confidence_threshold, self.class_threshold_thing, self.class_threshold_stuff,
                                             self.overlap_threshold, self.reorder_class_weight, self.reorder_mask_weight,
                                             seg.data_ptr(), segments.data_ptr(), ws.data_ptr(), ws.numel(), ops._stream(mp.device))
        _lib.check(rc, "axvs_panoptic_inference")
        return seg.view(out_shape), segments

    @torch.no_grad()
    def panoptic_mask_inference(self, mask_cls: torch.Tensor, mask_pred: torch.Tensor, mask_embedding: torch.Tensor):
        """mask_cls [N, C+1], mask_pred [N, T, H, W], mask_embedding [N, E] -> (panoptic_seg_mask, dic_cat_idemb)."""
        seg, segments = self.panoptic_segments(mask_cls, mask_pred)
        table = segments.cpu()                                            # the one host read (the reference does 3 per slot)
        n = int(table[0])
        rows = table[1:1 + 4 * n].view(n, 4).tolist()
        cat, _ = self._label_tables(mask_cls.shape[1] - 1, mask_pred.device)
        cat = cat.tolist()
        dic_cat_idemb: Dict[int, List[torch.Tensor]] = {}
        if n:
            slots = torch.tensor([r[0] for r in rows], dtype=torch.long, device=mask_embedding.device)
            emb = torch.nn.functional.normalize(mask_embedding[slots], p=2, dim=1)          # :549
            for k, (slot, label, is_thing, _) in enumerate(rows):
                if is_thing:
                    dic_cat_idemb.setdefault(cat[label], []).append(emb[k])
        return seg, dic_cat_idemb
