"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference modules by file path.

Only usable where ``/root/reference`` is mounted (the authoring container).  Never imported by the
product package, ``bench.py`` or any ``-m gpu`` test: the GPU box has no reference tree.  Used by
``oracle/make_golden.py`` (to generate ``tests/golden/*.npz``) and by the CPU-side tests that pin the
oracle restatement against the real reference (they skip when the tree is absent).

The reference files need ``detectron2`` / ``timm`` (absent here) only for registries, ``ShapeSpec``
and weight-init helpers, none of which touch the arithmetic of the hot path; small stand-ins are
installed in ``sys.modules`` before the import.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("AXVS_REFERENCE_ROOT", "/root/reference")
VK = os.path.join(REF_ROOT, "MaXTron_Video-kMaX")
WC = os.path.join(VK, "maxtron_deeplab/modeling/within_clip_tracking_module")
CC = os.path.join(VK, "maxtron_deeplab/modeling/cross_clip_tracking_module")


def available() -> bool:
    return os.path.isfile(os.path.join(WC, "temporal_attention.py"))


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _pkg(name: str, path: str | None = None):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    sys.modules[name] = m
    return m


def _install_stubs():
    import torch

    if "detectron2" not in sys.modules:
        d2 = _pkg("detectron2")
        cfgm = _pkg("detectron2.config")
        cfgm.configurable = lambda f=None, **kw: (f if f is not None else (lambda g: g))
        layers = _pkg("detectron2.layers")

        class ShapeSpec:  # noqa: D401 - stand-in for detectron2.layers.ShapeSpec
            def __init__(self, channels=None, height=None, width=None, stride=None):
                self.channels, self.height, self.width, self.stride = channels, height, width, stride

        layers.ShapeSpec = ShapeSpec
        modeling = _pkg("detectron2.modeling")

        class _Registry(dict):
            def __init__(self, name=""):
                super().__init__()
                self._name = name

            def register(self, obj=None):
                if obj is None:
                    return lambda o: self.register(o)
                self[obj.__name__] = obj
                return obj

        modeling.SEM_SEG_HEADS_REGISTRY = _Registry("SEM_SEG_HEADS")
        modeling.BACKBONE_REGISTRY = _Registry("BACKBONE")
        modeling.Backbone = torch.nn.Module
        modeling.ShapeSpec = ShapeSpec
        utils = _pkg("detectron2.utils")
        reg = _pkg("detectron2.utils.registry")
        reg.Registry = _Registry
        d2.config, d2.layers, d2.modeling, d2.utils = cfgm, layers, modeling, utils
    if "timm" not in sys.modules:
        timm = _pkg("timm")
        tm = _pkg("timm.models")
        tl = _pkg("timm.models.layers")

        class DropPath(torch.nn.Identity):
            def __init__(self, *a, **k):
                super().__init__()

        def trunc_normal_tf_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
            with torch.no_grad():
                torch.nn.init.trunc_normal_(t, 0.0, 1.0, a, b)
                t.mul_(std).add_(mean)
            return t

        tl.DropPath, tl.trunc_normal_tf_ = DropPath, trunc_normal_tf_
        timm.models, tm.layers = tm, tl


_cache: dict = {}


def temporal_attention():
    """`WC/temporal_attention.py` (TrajectoryAttention, TemporalEncoder, the two layer types)."""
    if "ta" not in _cache:
        _cache["ta"] = _load("_axvs_ref_temporal_attention", os.path.join(WC, "temporal_attention.py"))
    return _cache["ta"]


def pos_embeddings():
    """`WC/pos_embeddings.py` (PositionEmbeddingSine, PositionEmbeddingSine3D)."""
    if "pe" not in _cache:
        _cache["pe"] = _load("_axvs_ref_pos_embeddings", os.path.join(WC, "pos_embeddings.py"))
    return _cache["pe"]


def cross_clip():
    """`.../cross_clip_tracking_module/maxtron_cross_clip_tracking_module.py` with its two imports
    (`kmax_pixel_decoder.{get_norm,ConvBN}`, `maxtron_transformer_decoder.add_bias_towards_void`)."""
    if "cc" in _cache:
        return _cache["cc"]
    _install_stubs()
    # package skeletons that bypass the reference's __init__ files (they import detectron2.data)
    _pkg("kmax_deeplab", os.path.join(VK, "kmax_deeplab"))
    _pkg("kmax_deeplab.modeling", os.path.join(VK, "kmax_deeplab/modeling"))
    _pkg("kmax_deeplab.modeling.backbone", os.path.join(VK, "kmax_deeplab/modeling/backbone"))
    _pkg("kmax_deeplab.modeling.pixel_decoder", os.path.join(VK, "kmax_deeplab/modeling/pixel_decoder"))
    _pkg("maxtron_deeplab", os.path.join(VK, "maxtron_deeplab"))
    _pkg("maxtron_deeplab.modeling", os.path.join(VK, "maxtron_deeplab/modeling"))
    _pkg("maxtron_deeplab.modeling.transformer_decoder",
         os.path.join(VK, "maxtron_deeplab/modeling/transformer_decoder"))
    _load("kmax_deeplab.modeling.backbone.convnext",
          os.path.join(VK, "kmax_deeplab/modeling/backbone/convnext.py"))
    _load("kmax_deeplab.modeling.pixel_decoder.kmax_pixel_decoder",
          os.path.join(VK, "kmax_deeplab/modeling/pixel_decoder/kmax_pixel_decoder.py"))
    _load("maxtron_deeplab.modeling.transformer_decoder.maxtron_transformer_decoder",
          os.path.join(VK, "maxtron_deeplab/modeling/transformer_decoder/maxtron_transformer_decoder.py"))
    _cache["cc"] = _load("_axvs_ref_cross_clip",
                         os.path.join(CC, "maxtron_cross_clip_tracking_module.py"))
    return _cache["cc"]


def within_clip_module():
    """`WC/msdeformattn.py` (MSDeformAttn spatial layer, encoder, pixel-decoder plumbing) as a synthetic package so that its
    relative imports (`.ops.modules`, `.pos_embeddings`, `.temporal_attention`) resolve.  The compiled
    `MultiScaleDeformableAttention` extension is replaced by an empty module: `MSDeformAttn.forward` then falls into its own
    pure-PyTorch `ms_deform_attn_core_pytorch` branch (WC/ops/modules/ms_deform_attn.py:116-121) -- the CPU path of the reference."""
    if "wcm" in _cache:
        return _cache["wcm"]
    _install_stubs()
    if "MultiScaleDeformableAttention" not in sys.modules:
        sys.modules["MultiScaleDeformableAttention"] = types.ModuleType("MultiScaleDeformableAttention")
    pkg = "_axvs_ref_wcm"
    _pkg(pkg, WC)
    _pkg(pkg + ".ops", os.path.join(WC, "ops"))
    for sub in ("functions", "modules"):
        d = os.path.join(WC, "ops", sub)
        spec = importlib.util.spec_from_file_location(f"{pkg}.ops.{sub}", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"{pkg}.ops.{sub}"] = mod
        spec.loader.exec_module(mod)
    for name in ("pos_embeddings", "temporal_attention", "msdeformattn"):
        _load(f"{pkg}.{name}", os.path.join(WC, name + ".py"))
    _cache["wcm"] = sys.modules[pkg + ".msdeformattn"]
    return _cache["wcm"]


def wc_model():
    """`Vk/maxtron_deeplab/maxtron_wc_model.py` (MaXTronWCDeepLab: `panoptic_mask_inference`, the post-path tail).
    detectron2 is needed only for registries / type names at import time, and the two sibling modules (criterion, matcher: training
    only) are replaced by empty stand-ins."""
    if "wcmodel" in _cache:
        return _cache["wcmodel"]
    _install_stubs()
    import torch
    d2 = sys.modules["detectron2"]
    data = _pkg("detectron2.data")
    data.MetadataCatalog = types.SimpleNamespace(get=lambda name: types.SimpleNamespace())
    modeling = sys.modules["detectron2.modeling"]
    if not hasattr(modeling, "META_ARCH_REGISTRY"):
        modeling.META_ARCH_REGISTRY = type(modeling.SEM_SEG_HEADS_REGISTRY)("META_ARCH")
        modeling.build_backbone = modeling.build_sem_seg_head = lambda *a, **k: None
    bb = _pkg("detectron2.modeling.backbone")
    bb.Backbone = torch.nn.Module
    st = _pkg("detectron2.structures")
    st.ImageList = object
    mem = _pkg("detectron2.utils.memory")
    mem.retry_if_cuda_oom = lambda f: f
    d2.data, d2.structures = data, st
    pkg = "_axvs_ref_maxtron"
    _pkg(pkg, os.path.join(VK, "maxtron_deeplab"))
    _pkg(pkg + ".modeling", os.path.join(VK, "maxtron_deeplab/modeling"))
    crit = _pkg(pkg + ".modeling.wc_criterion")
    crit.MaXTronWCSetCriterion = object
    mat = _pkg(pkg + ".modeling.matcher")
    mat.VideoHungarianMatcher = object
    _cache["wcmodel"] = _load(pkg + ".maxtron_wc_model", os.path.join(VK, "maxtron_deeplab/maxtron_wc_model.py"))
    return _cache["wcmodel"]


# ------------------------------------------------------------------------------------------------ Tube-Link source slices
# The Tube-Link files import mmcv / mmdet / mmengine at module level (absent here), but the classes on the hot path are pure
# torch + einops.  They are taken as LINE SLICES of the unmodified files and exec'd in a namespace that provides only the names the
# slice uses: nothing is copied into the repository.
TL = os.path.join(REF_ROOT, "MaXTron_Tube-Link")
TL_PIXEL_DECODER = os.path.join(TL, "mmdet/models/plugins/msdeformattn_pixel_decoder.py")
TL_CC_HEAD = os.path.join(TL, "models/video/tube_link_vis/mask2former_video_cc_head.py")


def tl_available() -> bool:
    return os.path.isfile(TL_PIXEL_DECODER) and os.path.isfile(TL_CC_HEAD)


def _exec_slice(path: str, first: int, last: int, extra: dict | None = None) -> dict:
    """Execute lines [first, last] (1-based, inclusive) of `path` in a fresh namespace with torch / einops names."""
    import math
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from einops import rearrange
    with open(path) as fh:
        lines = fh.readlines()
    src = "".join(lines[first - 1:last])
    ns = {"math": math, "torch": torch, "nn": nn, "F": F, "Tensor": torch.Tensor, "rearrange": rearrange, "__name__": "tl_slice"}
    ns.update(extra or {})
    exec(compile(src, f"{path}:{first}-{last}", "exec"), ns)
    return ns


def tl_temporal_classes():
    """TrajectoryAttention, TemporalEncoder, TemporalAxialTrajectoryAttentionLayer of the Tube-Link pixel decoder (TL :640-791)."""
    ns = _exec_slice(TL_PIXEL_DECODER, 640, 791)
    return ns["TrajectoryAttention"], ns["TemporalEncoder"], ns["TemporalAxialTrajectoryAttentionLayer"]


def tl_cc_classes():
    """TrajectoryAttentionLayer and TrajectoryAttention of the Tube-Link cross-clip head (TL cc head :152-247)."""
    helper = _exec_slice(TL_CC_HEAD, 62, 71)                       # _get_activation_fn
    ns = _exec_slice(TL_CC_HEAD, 152, 247, {"_get_activation_fn": helper["_get_activation_fn"]})
    return ns["TrajectoryAttentionLayer"], ns["TrajectoryAttention"]
