"""Registry / config surface of the drop-ins ("what calls the path").

Video-kMaX selects the within-clip tracking module by NAME through Detectron2's `SEM_SEG_HEADS_REGISTRY`
(`cfg.MODEL.MAXTRON.WITHIN_CLIP_TRACKING_MODULE.NAME`, Vk/maxtron_deeplab/modeling/meta_arch/maxtron_deeplab_head.py:16-22) and builds it with
`from_config` (WC/maxtron_within_clip_tracking_module.py:44-63).  Tube-Link selects its attention plugin by `type` through MMCV's
`ATTENTION` registry (TL/mmdet/models/plugins/msdeformattn_pixel_decoder.py:393).  `register()` adds the B200 classes to those registries
when the frameworks are importable, under NEW names, so switching is a one-line config change and the reference classes stay available:

    MODEL.MAXTRON.WITHIN_CLIP_TRACKING_MODULE.NAME: "B200WithinClipTrackingModule"          # Video-kMaX yaml
    attn_cfgs=dict(type='B200MultiScaleDeformableAxialTrajectoryAttention', ...)           # Tube-Link python config

Neither framework is needed to import this module or to use the classes directly (they are plain nn.Modules).
"""
from __future__ import annotations

from typing import Dict

from torch import nn

from . import within_clip


def within_clip_kwargs_from_cfg(cfg, input_shape: Dict[str, object]) -> dict:
    """The reference's `WithinClipTrackingModule.from_config` (WC/maxtron_within_clip_tracking_module.py:44-63), key for key."""
    wc = cfg.MODEL.MAXTRON.WITHIN_CLIP_TRACKING_MODULE
    return {
        "input_shape": {k: v for k, v in input_shape.items() if k in wc.SPATIAL_IN_FEATURES},
        "transformer_dropout": wc.DROPOUT,
        "transformer_attn_drop": wc.ATTN_DROP,
        "transformer_nheads": wc.NHEADS,
        "transformer_dim_feedforward": wc.DIM_FEEDFORWARD,
        "transformer_num_stages": wc.NUM_STAGES,
        "transformer_spatial_layers": wc.SPATIAL_LAYERS,
        "transformer_temporal_layers": wc.TEMPORAL_LAYERS,
        "transformer_temporal_attn_type": wc.TEMPORAL_ATTN_TYPE,
        "transformer_conv_dims": wc.CONV_DIMS,
        "transformer_spatial_in_features": wc.SPATIAL_IN_FEATURES,
        "transformer_temporal_in_features": wc.TEMPORAL_IN_FEATURES,
        "num_clip_frames": cfg.INPUT.NUM_CLIP_FRAMES,
        "cross_clip_training": cfg.MODEL.MAXTRON.CROSS_CLIP_TRACKING_MODULE.ENABLE,
    }


class B200WithinClipTrackingModule(nn.Module):
    """Same constructor keywords, `from_config`, sub-module name (`within_clip_tracking_module`, so checkpoints load unchanged) and
    `forward_features` as the reference's registry class (WC/maxtron_within_clip_tracking_module.py:14-69)."""

    def __init__(self, input_shape, *, transformer_dropout, transformer_attn_drop, transformer_nheads, transformer_dim_feedforward,
                 transformer_num_stages, transformer_spatial_layers, transformer_temporal_layers, transformer_temporal_attn_type,
                 transformer_conv_dims, transformer_spatial_in_features, transformer_temporal_in_features, num_clip_frames, cross_clip_training):
        super().__init__()
        self.within_clip_tracking_module = within_clip.WithinClipTrackingModule(
            input_shape, transformer_dropout=transformer_dropout, transformer_attn_drop=transformer_attn_drop, transformer_nheads=transformer_nheads,
            transformer_dim_feedforward=transformer_dim_feedforward, transformer_num_stages=transformer_num_stages,
            transformer_spatial_layers=transformer_spatial_layers, transformer_temporal_layers=transformer_temporal_layers,
            transformer_temporal_attn_type=transformer_temporal_attn_type, conv_dims=transformer_conv_dims,
            transformer_spatial_in_features=transformer_spatial_in_features, transformer_temporal_in_features=transformer_temporal_in_features,
            num_clip_frames=num_clip_frames, cross_clip_training=cross_clip_training)

    @classmethod
    def from_config(cls, cfg, input_shape):
        return within_clip_kwargs_from_cfg(cfg, input_shape)

    def forward_features(self, features):
        out, h_attn, w_attn = self.within_clip_tracking_module.forward_features(features)      # :65-69
        for k in out:
            features[k] = out[k]
        return features, h_attn, w_attn


def register() -> Dict[str, str]:
    """Add the drop-ins to the host frameworks' registries; returns {framework: outcome}.  Safe to call more than once."""
    done: Dict[str, str] = {}
    try:
        from detectron2.config import configurable
        from detectron2.modeling import SEM_SEG_HEADS_REGISTRY
        name = "B200WithinClipTrackingModule"
        if name not in SEM_SEG_HEADS_REGISTRY:
            cls = type(name, (B200WithinClipTrackingModule,), {})
            cls.__init__ = configurable(B200WithinClipTrackingModule.__init__)          # Detectron2 calls cls(cfg, input_shape)
            cls.from_config = classmethod(lambda c, cfg, input_shape: within_clip_kwargs_from_cfg(cfg, input_shape))
            SEM_SEG_HEADS_REGISTRY.register(cls)
        done["detectron2"] = f"SEM_SEG_HEADS_REGISTRY['{name}']"
    except ImportError:
        done["detectron2"] = "not importable (classes usable directly)"
    try:
        from mmcv.cnn.bricks.registry import ATTENTION
        from .tube_link import MultiScaleDeformableAxialTrajectoryAttention
        name = "B200MultiScaleDeformableAxialTrajectoryAttention"
        if name not in ATTENTION.module_dict:
            ATTENTION.register_module(name=name, module=MultiScaleDeformableAxialTrajectoryAttention)
        done["mmcv"] = f"ATTENTION['{name}']"
    except ImportError:
        done["mmcv"] = "not importable (classes usable directly)"
    return done
