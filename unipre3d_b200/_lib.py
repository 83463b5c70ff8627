"""ctypes binding of libunipre3d_b200.so (the C ABI declared in include/up3d.h).

There is NO CPU fallback: if the shared library is missing and cannot be built, importing this
module raises; if a tensor is not on a CUDA device the host wrappers raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libunipre3d_b200.so")


class RasterDesc(C.Structure):
    """struct up3d_raster_desc (include/up3d.h)."""
    _fields_ = [("n_sets", C.c_int32), ("n_views", C.c_int32), ("n_gaussians", C.c_int32), ("n_records", C.c_int32),
                ("max_set_size", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("sh_degree", C.c_int32),
                ("sh_coeffs", C.c_int32), ("antialiasing", C.c_int32), ("tanfovx", C.c_float), ("tanfovy", C.c_float),
                ("scale_modifier", C.c_float),
                ("set_offsets", C.c_void_p), ("set_view_start", C.c_void_p), ("view_set", C.c_void_p),
                ("view_rec_start", C.c_void_p)]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        # in-tree build (nvcc cross-compiles without a GPU); fails loudly when nvcc is absent
        from .csrc.build import build_library
        build_library()
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    D = C.POINTER(RasterDesc)
    sigs = {
        "up3d_last_error": (C.c_char_p, []),
        "up3d_version": (i32, []),
        "up3d_fps": (i32, [i32, i32, i32, vp, vp, vp, vp]),
        "up3d_fps_max_resident_points": (i32, []),
        "up3d_ball_query": (i32, [i32, i32, i32, f32, i32, vp, vp, vp, vp]),
        "up3d_group_points": (i32, [i32, i32, i32, i32, i32, vp, vp, vp, vp]),
        "up3d_group_points_grad": (i32, [i32, i32, i32, i32, i32, vp, vp, vp, vp]),
        "up3d_gather_points": (i32, [i32, i32, i32, i32, vp, vp, vp, vp]),
        "up3d_gather_points_grad": (i32, [i32, i32, i32, i32, vp, vp, vp, vp]),
        "up3d_subsample_group": (i32, [i32, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp]),
        "up3d_knn": (i32, [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
        "up3d_raster_state_bytes": (C.c_size_t, [D]),
        "up3d_raster_scratch_bytes": (C.c_size_t, [D]),
        "up3d_raster_forward": (i32, [D] + [vp] * 16),
        "up3d_raster_backward": (i32, [D] + [vp] * 21),
        "up3d_raster_debug_state": (i32, [D] + [vp] * 11),
        "up3d_raster_debug_tile_lists": (i32, [D] + [vp] * 5),
        "up3d_raster_debug_bins": (i32, [D] + [vp] * 4),
        "up3d_focal_l2_loss": (i32, [i64, i32, i32, vp, vp, vp, f32, f32, vp, vp, vp]),
        "up3d_focal_l2_loss_strided": (i32, [i64, i32, i32, vp, vp, i32, i64, i64, vp, f32, f32, vp, vp, vp]),
        "up3d_raster_timing_enable": (i32, [i32]),
        "up3d_raster_timing_read": (i32, [C.POINTER(C.c_float)]),
        "up3d_ln_fwd": (i32, [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp, vp]),
        "up3d_ln_bwd": (i32, [i32, i32, i32, i32] + [vp] * 14),
        "up3d_gelu_fwd": (i32, [i32, i64, vp, vp, vp]),
        "up3d_gelu_bwd": (i32, [i32, i32, i32, vp, vp, vp, vp, vp]),
        "up3d_scale_cast_colsum": (i32, [i32, i32, i32, i32, vp, vp, vp, vp, vp]),
        "up3d_stem_group_stats": (i32, [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
        "up3d_pn_stats_tile_rows": (i32, []),
        "up3d_pn_conv1_stats": (i32, [i32, i32, vp, vp, vp, vp, vp]),
        "up3d_bn_reduce_sums": (i32, [i32, i32, i32, vp, vp, vp]),
        "up3d_pn_conv1_bn_relu": (i32, [i32, i32, i32, vp, vp, vp, vp, vp, vp]),
        "up3d_pn_conv1_bwd": (i32, [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, C.c_double, vp, i32, vp, vp, vp]),
        "up3d_bn_reduce_finalize": (i32, [i32, i32, vp, i32, i64, vp, vp, f32, f32, vp, vp, vp, vp, vp, vp]),
        "up3d_gbn_stats": (i32, [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
        "up3d_gbn_apply_relu": (i32, [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
        "up3d_gbn_bwd_reduce": (i32, [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
        "up3d_gbn_bwd_apply": (i32, [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, C.c_double, vp, vp, vp]),
        "up3d_group_max": (i32, [i32, i32, i32, i32, vp, vp, vp, vp]),
        "up3d_group_max_scatter": (i32, [i32, i32, i32, i32, vp, vp, vp, vp]),
        "up3d_group_combine": (i32, [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
        "up3d_attn_max_len": (i32, []),
        "up3d_attn_fwd": (i32, [i32, i32, i32, i32, f32, vp, vp, vp, vp]),
        "up3d_attn_bwd": (i32, [i32, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp]),
        "up3d_splat_head_fwd": (i32, [i32, i32, i32, vp, vp, f32, i32, vp, vp, vp, vp, vp, vp, vp]),
        "up3d_splat_head_bwd": (i32, [i32, i32, i32, vp, vp, f32, i32, vp, vp, vp, vp, vp, vp, vp]),
        "up3d_fusion_project": (i32, [i32] * 6 + [f32] * 5 + [vp] * 10),
        "up3d_zorder_keys": (i32, [i64, i32, i32, vp, vp, vp, vp]),
        "up3d_hilbert_keys": (i32, [i64, i32, i32, vp, vp, vp, vp]),
        "up3d_adamw_chunk_elems": (i32, []),
        "up3d_adamw_step": (i32, [i32, i32] + [vp] * 10 + [f32] * 6 + [vp, vp]),
        "up3d_adamw_apply": (i32, [i32, i32] + [vp] * 10 + [f32] * 6 + [vp, vp]),
        "up3d_grad_sumsq": (i32, [i32, i32] + [vp] * 4 + [f32, vp, vp]),
        "up3d_set_pdl": (i32, [i32]),
        "up3d_set_group_tile_staging": (i32, [i32]),
        "up3d_sparse_subm_rulebook": (i32, [i32, i32, vp, vp, vp, vp]),
        "up3d_sparse_conv": (i32, [i32, i32, i32, i32, vp, vp, vp, vp, vp]),
        "up3d_sparse_conv_wgrad": (i32, [i32, i32, i32, i32, vp, vp, vp, vp, vp]),
        "up3d_tc_linear": (i32, [i32, i32, i32, vp, vp, i32, vp, i32, vp, vp, vp, i32, i32, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
EXPORTED = ["up3d_last_error", "up3d_version", "up3d_fps", "up3d_fps_max_resident_points", "up3d_ball_query",
            "up3d_group_points", "up3d_group_points_grad", "up3d_gather_points", "up3d_gather_points_grad",
            "up3d_subsample_group", "up3d_knn", "up3d_raster_state_bytes", "up3d_raster_scratch_bytes", "up3d_raster_forward",
            "up3d_raster_backward", "up3d_raster_debug_state", "up3d_raster_debug_tile_lists", "up3d_raster_debug_bins", "up3d_focal_l2_loss",
            "up3d_focal_l2_loss_strided",
            "up3d_raster_timing_enable", "up3d_raster_timing_read", "up3d_ln_fwd", "up3d_ln_bwd", "up3d_gelu_fwd",
            "up3d_gelu_bwd", "up3d_scale_cast_colsum", "up3d_adamw_chunk_elems", "up3d_adamw_step", "up3d_adamw_apply",
            "up3d_grad_sumsq", "up3d_stem_group_stats", "up3d_pn_conv1_stats", "up3d_pn_stats_tile_rows",
            "up3d_bn_reduce_sums", "up3d_pn_conv1_bn_relu", "up3d_pn_conv1_bwd",
            "up3d_bn_reduce_finalize", "up3d_gbn_stats", "up3d_gbn_apply_relu", "up3d_gbn_bwd_reduce", "up3d_gbn_bwd_apply",
            "up3d_group_max", "up3d_group_max_scatter", "up3d_group_combine", "up3d_attn_max_len", "up3d_attn_fwd",
            "up3d_attn_bwd", "up3d_splat_head_fwd", "up3d_splat_head_bwd", "up3d_fusion_project",
            "up3d_zorder_keys", "up3d_hilbert_keys", "up3d_tc_linear", "up3d_set_pdl", "up3d_set_group_tile_staging", "up3d_sparse_subm_rulebook",
            "up3d_sparse_conv", "up3d_sparse_conv_wgrad"]

# kernels launched by this process through the C ABI (bench.py reports it as gpu_launches)
launch_count = 0


def check(status: int, launches: int = 0) -> None:
    """Non-zero status -> RuntimeError carrying up3d_last_error() (the external rasterizer the reference
    uses raises std::runtime_error -> RuntimeError; we keep that error type)."""
    global launch_count
    if status != 0:
        raise RuntimeError(lib.up3d_last_error().decode("utf-8", "replace"))
    launch_count += launches


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("unipre3d_b200: tensors must live on a CUDA device (there is no CPU fallback); "
                               f"got device {t.device}")
