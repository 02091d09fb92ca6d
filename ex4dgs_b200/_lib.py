"""ctypes loader for libex4dgs_raster.so (the C ABI declared in include/ex4dgs_raster.h).

The product path has NO fallback: if the CUDA library is missing or fails to load, importing a
symbol raises immediately (a silent CPU/eager fallback would void every parity claim).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EX4DGS_LIB") or os.path.join(HERE, "libex4dgs_raster.so")   # EX4DGS_LIB: tuning experiments only

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)

FLAG_TILE_CULL = 1
FLAG_SH_SEGMENTED = 2
FLAG_NO_HOST_WAIT = 4

_F = C.c_float
_P = C.c_void_p
_I = C.c_int
_D = C.c_double


class ShSegments(C.Structure):
    """ex4dgs_sh_segments (include/ex4dgs_raster.h)."""
    _fields_ = [("n_static", C.c_int), ("dc_static", C.c_void_p), ("rest_static", C.c_void_p),
                ("dc_dynamic", C.c_void_p), ("rest_dynamic", C.c_void_p)]


class RAdamTensor(C.Structure):
    """ex4dgs_radam_tensor (include/ex4dgs_raster.h)."""
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_size_t), ("lr", C.c_double), ("step", C.c_longlong)]


class GatherJob(C.Structure):
    """ex4dgs_gather_job (include/ex4dgs_raster.h)."""
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("dst", C.c_void_p), ("index", C.c_void_p),
                ("row_bytes", C.c_size_t), ("n_a", C.c_longlong), ("n_out", C.c_longlong)]


class StatsArrays(C.Structure):
    """ex4dgs_stats_arrays (include/ex4dgs_raster.h)."""
    _fields_ = [(n, C.c_void_p) for n in ("max_radii2D", "min_radii2D", "xyz_gradient_accum", "denom", "error_accum",
                                          "error_min", "error_min_timestamp", "ssim_error_accum", "error_denom")]


class ArrayDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("buffer", C.c_int), ("offset", C.c_size_t),
                ("elem_size", C.c_size_t), ("count", C.c_size_t)]


SIGNATURES = {
    # name: (restype, argtypes)
    "ex4dgs_abi_version": (_I, []),
    "ex4dgs_last_error": (C.c_char_p, []),
    "ex4dgs_last_inexact_thresholds": (C.c_uint, []),
    "ex4dgs_forward_geometry": (None, [C.POINTER(_I), C.POINTER(_I)]),
    "ex4dgs_set_capacity_hint": (None, [_I]),
    "ex4dgs_geometry_bytes": (C.c_size_t, [_I]),
    "ex4dgs_binning_bytes": (C.c_size_t, [_I]),
    "ex4dgs_image_bytes": (C.c_size_t, [_I, _I]),
    "ex4dgs_describe_buffers": (_I, [_I, _I, _I, _I, C.POINTER(ArrayDesc), _I]),
    "ex4dgs_forward": (_I, [ALLOC_FN, _P, ALLOC_FN, _P, ALLOC_FN, _P,
                            _I, _I, _I,          # P D M
                            _P, _I, _I,          # background width height
                            _P, _P, _P, _P,      # means3D dir3D shs colors_precomp
                            _P, _P, _F, _P,      # opacities scales scale_modifier rotations
                            _P, _P, _P, _P,      # cov3D_precomp viewmatrix projmatrix cam_pos
                            _F, _F, _F, _P, _I,  # tan_fovx tan_fovy kernel_size subpixel_offset prefiltered
                            _P, _F, _F, _P, _P, _P,  # out_color min_depth max_depth out_depth out_acc out_flow
                            _P, _P, _I, C.c_uint, _P]),  # out_idx radii debug flags stream
    "ex4dgs_backward": (_I, [_I, _I, _I, _I,     # P D M R
                             _P, _I, _I,         # background width height
                             _P, _P, _P,         # means3D shs colors_precomp
                             _P, _F, _P,         # scales scale_modifier rotations
                             _P, _P, _F, _F,     # acc_depth acc min_depth max_depth
                             _P, _P, _P, _P,     # cov3D_precomp viewmatrix projmatrix campos
                             _F, _F, _F, _P, _P,  # tan_fovx tan_fovy kernel_size subpixel_offset radii
                             _P, _P, _P,         # geom binning image buffers
                             _P, _P, _P, _P,     # dL_dpix dL_ddepth dL_dflow dL_dacc
                             _P, _P, _P, _P, _P,  # dL_dmean2D dL_dopacity dL_dcolor dL_dmean3D dL_dcov3D
                             _P, _P, _P, _P,     # dL_dsh dL_dscale dL_drot dL_ddir
                             _I, C.c_uint, _P]),  # debug flags stream
    "ex4dgs_profile_enable": (None, [_I]),
    "ex4dgs_profile_read": (_I, [C.POINTER(C.c_double), C.POINTER(_I), C.POINTER(_I)]),
    "ex4dgs_launch_count": (C.c_ulonglong, []),
    "ex4dgs_mark_visible": (_I, [_I, _P, _P, _P, _F, _F, _P, _P]),
    "ex4dgs_frontend_forward": (_I, [_I, _I, _I,
                                     _P, _P, _P, _P, _P,
                                     _P, _P, _P, _P, _P, _P,
                                     _D, _D, _D, _D, _D,
                                     _P, _P, _P, _P, _P]),
    "ex4dgs_frontend_backward": (_I, [_I, _I, _I,
                                      _P, _P, _P, _P, _P, _P, _P,
                                      _D, _D, _D, _D, _D,
                                      _P, _P, _P, _P,
                                      _P, _P, _P, _P, _P,
                                      _P, _P, _P, _P, _P, _P,
                                      _P]),
    "ex4dgs_radam_step": (_I, [C.POINTER(RAdamTensor), _I, _D, _D, _D, _D, _P]),
    "ex4dgs_radam_step_ex": (_I, [C.POINTER(RAdamTensor), _I, _D, _D, _D, _D, C.c_uint, C.c_uint, _P, _P]),
    "ex4dgs_gather_rows": (_I, [C.POINTER(GatherJob), _I, _P]),
    "ex4dgs_radam_scalars": (_I, [_D, C.c_longlong, _D, _D, C.POINTER(_F), C.POINTER(_F), C.POINTER(_I)]),
    "ex4dgs_iteration_stats": (_I, [_I, _I, _P, _P, _P, _F, _I, C.POINTER(StatsArrays), C.POINTER(StatsArrays), _P]),
    "ex4dgs_regularizer_scratch_bytes": (C.c_size_t, []),
    "ex4dgs_regularizers": (_I, [_I, _I, _I, _P, _P, _F, _F, _P, _P, _I, _P, _I, _P, _P, _P]),
    "ex4dgs_l1_scratch_bytes": (C.c_size_t, []),
    "ex4dgs_l1_forward": (_I, [C.c_size_t, _P, _P, _P, _P, _P]),
    "ex4dgs_l1_backward": (_I, [C.c_size_t, _P, _P, _P, _P, _P]),
    "ex4dgs_loss_scratch_bytes": (C.c_size_t, [_I, _I]),
    "ex4dgs_loss_forward": (_I, [_I, _I, _P, _P, _F, _P, _P, _P, _P, _P]),
    "ex4dgs_loss_backward": (_I, [_I, _I, _P, _P, _F, _P, _P, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach the prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "ex4dgs_b200: %s is missing - build it with `python -m ex4dgs_b200.build` "
            "(there is deliberately no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.ex4dgs_abi_version() != 1:
        raise ImportError("ex4dgs_b200: ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load().ex4dgs_last_error().decode("utf-8", "replace")


def describe_buffers(P: int, R: int, W: int, H: int):
    lib = load()
    arr = (ArrayDesc * 32)()
    n = lib.ex4dgs_describe_buffers(P, R, W, H, arr, 32)
    return {arr[i].name.decode(): (arr[i].buffer, arr[i].offset, arr[i].elem_size, arr[i].count) for i in range(n)}
