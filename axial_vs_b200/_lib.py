"""ctypes binding of libaxvs.so (the C ABI declared in include/axvs.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised
(mirrors `ModuleNotFoundError` on the reference's native op, WC/ops/functions/ms_deform_attn_func.py:21-29,
and `AT_ERROR` -> RuntimeError, WC/ops/src/ms_deform_attn.h:43).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AXVS_LIB", os.path.join(_HERE, "libaxvs.so"))   # AXVS_LIB: debug/profiling builds

AXIS_NONE, AXIS_H, AXIS_W = 0, 1, 2


class TaWeights(Structure):
    _fields_ = [("w_qkv", c_void_p), ("b_qkv", c_void_p), ("w_pq", c_void_p), ("b_pq", c_void_p),
                ("w_pkv", c_void_p), ("b_pkv", c_void_p), ("w_qkv_u", c_void_p), ("w_pq_u", c_void_p), ("w_pkv_u", c_void_p), ("w_proj_u", c_void_p), ("w_proj", c_void_p), ("b_proj", c_void_p)]


class MsdaWeights(Structure):
    """struct axvs_msda_weights."""
    _fields_ = [("w_value", c_void_p), ("b_value", c_void_p), ("w_oa", c_void_p), ("b_oa", c_void_p), ("w_out", c_void_p), ("b_out", c_void_p),
                ("ln1_g", c_void_p), ("ln1_b", c_void_p), ("w_ffn1", c_void_p), ("b_ffn1", c_void_p), ("w_ffn2", c_void_p), ("b_ffn2", c_void_p),
                ("w_ffn1_u", c_void_p), ("w_ffn2_u", c_void_p), ("w_ffn1_n", c_void_p), ("ln2_g", c_void_p), ("ln2_b", c_void_p),
                ("w_front_u", c_void_p), ("b_front", c_void_p), ("w_out_u", c_void_p), ("d_ffn", c_int), ("n_levels", c_int), ("n_points", c_int)]


class KmaxAxialWeights(Structure):
    """struct axvs_kmax_axial_weights."""
    _fields_ = [("w_qkv", c_void_p), ("b_qkv", c_void_p), ("emb_q", c_void_p), ("emb_k", c_void_p), ("emb_v", c_void_p),
                ("sim_s", c_void_p), ("sim_t", c_void_p), ("out_s", c_void_p), ("out_t", c_void_p), ("heads", c_int), ("dk", c_int), ("dv", c_int)]


class AsppWeights(Structure):
    _fields_ = [("w_conv", c_void_p * 3), ("b_conv", c_void_p * 3), ("dilation", c_int * 3), ("w_proj", c_void_p),
                ("lncf_g", c_void_p), ("lncf_b", c_void_p), ("ln_g", c_void_p), ("ln_b", c_void_p), ("split", c_int)]


class LayerWeights(Structure):
    _fields_ = [("attn_h", TaWeights), ("attn_w", TaWeights), ("ln1_g", c_void_p), ("ln1_b", c_void_p),
                ("w_ffn1", c_void_p), ("b_ffn1", c_void_p), ("w_ffn2", c_void_p), ("b_ffn2", c_void_p),
                ("w_ffn1_u", c_void_p), ("w_ffn2_u", c_void_p), ("ln2_g", c_void_p), ("ln2_b", c_void_p), ("d_ffn", c_int), ("w_ffn1_n", c_void_p)]


# name -> (restype, argtypes); every symbol include/axvs.h declares
SIGNATURES = {
    "axvs_version": (c_int, []),
    "axvs_last_error": (c_char_p, []),
    "axvs_build_id": (c_char_p, []),
    "axvs_set_fusion": (c_int, [c_int]),
    "axvs_set_pair_mode": (c_int, [c_int]),
    "axvs_set_attn_core": (c_int, [c_int]),
    "axvs_packed_weight_bytes": (c_size_t, [c_int, c_int]),
    "axvs_pack_weight": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "axvs_pack_weight_units": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "axvs_linear": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_int, c_void_p,
                            c_int, c_int, c_void_p, c_void_p]),
    "axvs_spatial_attention": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "axvs_traj_attn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "axvs_traj_attn_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, POINTER(TaWeights), c_int, c_int, c_int,
                                   c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "axvs_traj_attn_maps": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(TaWeights), c_int, c_int, c_int, c_int, c_int,
                                    c_void_p, c_size_t, c_void_p]),
    "axvs_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p]),
    "axvs_ffn_workspace_bytes": (c_size_t, [c_int, c_int]),
    "axvs_ln_ffn_fwd": (c_int, [c_void_p, c_void_p, POINTER(LayerWeights), c_int, c_void_p, c_size_t, c_void_p]),
    "axvs_layer_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "axvs_axial_layer_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, POINTER(LayerWeights), c_int, c_int, c_int, c_int,
                                     c_int, c_void_p, c_size_t, c_void_p]),
    "axvs_cast_bf16": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "axvs_cc_aspp_workspace_bytes": (c_size_t, [c_int]),
    "axvs_cc_aspp_fwd": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(AsppWeights), c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "axvs_cc_class_pool": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_void_p]),
    "axvs_query_self_attn": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "axvs_cm_to_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "axvs_rows_to_cm": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "axvs_dwconv5": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "axvs_add_act": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p]),
    "axvs_masked_mha_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "axvs_masked_mha_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "axvs_frame_attn_f32_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "axvs_frame_attn_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "axvs_kmeans_update_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "axvs_kmeans_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "axvs_proj_workspace_bytes": (c_size_t, [c_int]),
    "axvs_input_proj_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "axvs_output_proj_fwd": (c_int, [c_void_p, ctypes.c_longlong, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "axvs_msda_layer_workspace_bytes": (c_size_t, [c_int, c_int]),
    "axvs_msda_layer_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, POINTER(c_int), c_void_p, POINTER(MsdaWeights), c_int, c_int, c_void_p, c_size_t,
                                    c_void_p]),
    "axvs_set_kmax_tensor_cores": (c_int, [c_int]),
    "axvs_kmax_axial_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "axvs_kmax_axial_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(KmaxAxialWeights), c_void_p, c_int, c_void_p, c_size_t,
                                    c_void_p]),
    "axvs_panoptic_workspace_bytes": (c_size_t, [c_int, ctypes.c_longlong]),
    "axvs_panoptic_inference": (c_int, [c_void_p, c_void_p, c_int, c_int, ctypes.c_longlong, c_void_p, c_void_p, c_int, c_float, c_float, c_float,
                                        c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "axvs_mask_einsum": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "axvs_msda_sample_workspace_bytes": (c_size_t, [c_int]),
    "axvs_msda_sample_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                     c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "axvs_lsap": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "axvs_match_chain_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "axvs_match_chain": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "axvs_mask_einsum_f32": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "axvs_linear_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_float, c_int, c_void_p, c_int, c_int,
                                c_void_p]),
    "axvs_pos3d": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "axvs_profile_enable": (c_int, [c_int]),
    "axvs_profile_num_classes": (c_int, []),
    "axvs_profile_class_name": (c_char_p, [c_int]),
    "axvs_profile_read": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libaxvs.so (built by `__graft_entry__.build()`); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"axial_vs_b200: native library not found at {LIB_PATH}. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU/PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().axvs_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
