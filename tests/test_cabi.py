"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ex4dgs_raster.h declares, and the Python surface mirrors the reference's L1 API
(names, field order, exceptions) - no compute calls, no GPU."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    from ex4dgs_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ex4dgs_raster.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ex4dgs_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ex4dgs_alloc_fn"}
    assert len(declared) >= 10
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    # and the ctypes prototypes cover exactly the declared set
    assert set(_lib.SIGNATURES) == declared


def test_abi_version_and_error_string(built):
    from ex4dgs_b200 import _lib
    lib = _lib.load()
    assert lib.ex4dgs_abi_version() == 1
    assert isinstance(_lib.last_error(), str)


def test_forward_rejects_bad_arguments_before_touching_cuda(built):
    from ex4dgs_b200 import _lib
    lib = _lib.load()
    cb = _lib.ALLOC_FN(lambda u, n: None)
    args = [cb, None, cb, None, cb, None, -1, 3, 16, None, 16, 16] + [None] * 4 + [None, None, 1.0, None] + [None] * 4 + \
           [1.0, 1.0, 0.1, None, 0] + [None, 0.2, 100.0, None, None, None] + [None, None, 0, 0, None]
    rc = lib.ex4dgs_forward(*args)
    assert rc == -1 and "bad sizes" in _lib.last_error()


def test_python_surface_matches_reference_names():
    import diff_gaussian_rasterization_df as m
    fields = ("image_height", "image_width", "tanfovx", "tanfovy", "kernel_size", "subpixel_offset", "bg",
              "scale_modifier", "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered", "min_depth",
              "max_depth", "debug")
    assert m.GaussianRasterizationSettings._fields == fields          # __init__.py:180-196
    sig = inspect.signature(m.GaussianRasterizer.forward)
    assert list(sig.parameters) == ["self", "means3D", "means2D", "dir3D", "opacities", "shs", "colors_precomp",
                                    "scales", "rotations", "cov3D_precomp"]     # __init__.py:215
    assert list(inspect.signature(m.rasterize_gaussians).parameters) == [
        "means3D", "means2D", "dir3D", "sh", "colors_precomp", "opacities", "scales", "rotations", "cov3Ds_precomp",
        "raster_settings"]                                              # __init__.py:22-33
    assert hasattr(m.GaussianRasterizer, "markVisible")


def _settings(mod):
    z = torch.zeros
    return mod.GaussianRasterizationSettings(16, 16, 0.5, 0.5, 0.1, z(16, 16, 2), z(3), 1.0, torch.eye(4), torch.eye(4), 3,
                                             z(3), False, 0.2, 100.0, False)


def test_argument_validation_raises_like_reference():
    """__init__.py:219-223: exactly one of shs/colors_precomp, scales+rotations xor cov3D_precomp."""
    import diff_gaussian_rasterization_df as m
    r = m.GaussianRasterizer(_settings(m))
    P = 4
    a = dict(means3D=torch.zeros(P, 3), means2D=torch.zeros(P, 3), dir3D=torch.zeros(P, 3), opacities=torch.ones(P, 1))
    with pytest.raises(Exception, match="excatly one of either SHs"):
        r(**a, scales=torch.ones(P, 3), rotations=torch.ones(P, 4))
    with pytest.raises(Exception, match="excatly one of either SHs"):
        r(**a, shs=torch.zeros(P, 16, 3), colors_precomp=torch.zeros(P, 3), scales=torch.ones(P, 3), rotations=torch.ones(P, 4))
    with pytest.raises(Exception, match="scale/rotation pair"):
        r(**a, shs=torch.zeros(P, 16, 3))
    with pytest.raises(Exception, match="scale/rotation pair"):
        r(**a, shs=torch.zeros(P, 16, 3), scales=torch.ones(P, 3), rotations=torch.ones(P, 4), cov3D_precomp=torch.ones(P, 6))


def test_no_cpu_fallback(built):
    """The product must fail loudly on CPU tensors (the reference is CUDA-only too)."""
    import diff_gaussian_rasterization_df as m
    r = m.GaussianRasterizer(_settings(m))
    P = 4
    with pytest.raises(RuntimeError, match="CUDA-only"):
        r(means3D=torch.zeros(P, 3), means2D=torch.zeros(P, 3), dir3D=torch.zeros(P, 3), opacities=torch.ones(P, 1),
          shs=torch.zeros(P, 16, 3), scales=torch.ones(P, 3), rotations=torch.ones(P, 4))
    with pytest.raises(RuntimeError, match="num_points, 3"):
        r(means3D=torch.zeros(P, 4), means2D=torch.zeros(P, 3), dir3D=torch.zeros(P, 3), opacities=torch.ones(P, 1),
          shs=torch.zeros(P, 16, 3), scales=torch.ones(P, 3), rotations=torch.ones(P, 4))


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "ex4dgs_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "liboracle" not in src and "cpu_raster" not in src, f


def test_training_side_entry_points_reject_bad_arguments_before_touching_cuda(built):
    """ex4dgs_iteration_stats / ex4dgs_regularizers / ex4dgs_l1_* / ex4dgs_radam_step_ex validate their arguments on
    the host and report through ex4dgs_last_error (no GPU needed to see that)."""
    from ex4dgs_b200 import _lib
    lib = _lib.load()
    st = _lib.StatsArrays()
    assert lib.ex4dgs_iteration_stats(-1, 0, None, None, None, 0.0, 1, ctypes.byref(st), ctypes.byref(st), None) < 0
    assert "negative" in _lib.last_error()
    assert lib.ex4dgs_iteration_stats(0, 0, None, None, None, 0.0, 1, None, None, None) == 0          # empty model: nothing to do
    assert lib.ex4dgs_iteration_stats(4, 0, None, None, None, 0.0, 1, ctypes.byref(st), None, None) < 0
    assert "required" in _lib.last_error()
    buf = (ctypes.c_int * 4)()
    fbuf = (ctypes.c_float * 12)()
    p, f = ctypes.addressof(buf), ctypes.addressof(fbuf)
    assert lib.ex4dgs_iteration_stats(4, 0, p, f, None, 0.0, 1, ctypes.byref(st), None, None) < 0     # densify needs the arrays
    assert "NULL" in _lib.last_error()
    assert lib.ex4dgs_regularizers(4, 0, 0, f, None, 1e-4, 0.0, None, None, 0, None, 0, None, None, None) < 0
    assert "required" in _lib.last_error()
    assert lib.ex4dgs_regularizers(-1, 0, 0, None, None, 0.0, 0.0, None, None, 0, None, 0, f, f, None) < 0
    assert lib.ex4dgs_l1_forward(0, f, f, f, f, None) < 0 and "bad arguments" in _lib.last_error()
    assert lib.ex4dgs_l1_backward(12, f, f, None, f, None) < 0
    t = (_lib.RAdamTensor * 1)()
    t[0].param = t[0].grad = t[0].exp_avg = t[0].exp_avg_sq = f
    t[0].numel, t[0].lr, t[0].step = 12, 1e-3, 1
    assert lib.ex4dgs_radam_step_ex(t, 1, 0.9, 0.999, 1e-8, 1.0, 1, 0, None, None) < 0                # check mask without flags
    assert "nan_flags" in _lib.last_error()
    assert lib.ex4dgs_radam_step_ex(t, 33, 0.9, 0.999, 1e-8, 1.0, 0, 0, None, None) < 0
    t[0].step = 0
    assert lib.ex4dgs_radam_step_ex(t, 1, 0.9, 0.999, 1e-8, 1.0, 0, 0, None, None) < 0 and "step" in _lib.last_error()
    assert lib.ex4dgs_regularizer_scratch_bytes() >= 2 * 8 and lib.ex4dgs_l1_scratch_bytes() >= 8


def test_gather_rows_rejects_bad_jobs_before_touching_cuda(built):
    """ex4dgs_gather_rows (densification / pruning row gathers): host-side validation of the job table."""
    from ex4dgs_b200 import _lib
    lib = _lib.load()
    fbuf = (ctypes.c_float * 12)()
    f = ctypes.addressof(fbuf)
    j = (_lib.GatherJob * 1)()
    assert lib.ex4dgs_gather_rows(j, 0, None) == 0                                           # empty table
    assert lib.ex4dgs_gather_rows(j, 65, None) < 0 and "outside" in _lib.last_error()
    j[0].a, j[0].dst, j[0].row_bytes, j[0].n_a, j[0].n_out = f, f, 12, 0, 0
    assert lib.ex4dgs_gather_rows(j, 1, None) == 0                                           # no output rows: nothing to do
    j[0].n_a, j[0].n_out = 3, 2
    assert lib.ex4dgs_gather_rows(j, 1, None) < 0 and "n_a" in _lib.last_error()
    j[0].n_a, j[0].n_out, j[0].row_bytes = 1, 1, 6
    assert lib.ex4dgs_gather_rows(j, 1, None) < 0 and "multiple of 4" in _lib.last_error()
    j[0].row_bytes, j[0].dst = 12, None
    assert lib.ex4dgs_gather_rows(j, 1, None) < 0 and "NULL" in _lib.last_error()
    j[0].dst, j[0].a = f, f + 2
    assert lib.ex4dgs_gather_rows(j, 1, None) < 0 and "aligned" in _lib.last_error()


def test_describe_buffers_switches_to_32_bit_tile_keys_beyond_65535_tiles(built):
    """Host-only layout query: 16-bit tile keys while every tile id and the all-ones dump key fit, 32-bit keys for larger
    images; ex4dgs_binning_bytes is an upper bound for both."""
    from ex4dgs_b200 import _lib
    lib = _lib.load()
    small = _lib.describe_buffers(1000, 5000, 1352, 1014)
    large = _lib.describe_buffers(1000, 5000, 4112, 4112)         # 257 x 257 = 66 049 tiles
    edge = _lib.describe_buffers(1000, 5000, 4080, 4112)          # 255 x 257 = 65 535 tiles: still 16-bit
    assert small["tile_sorted"][2] == 2 and edge["tile_sorted"][2] == 2 and large["tile_sorted"][2] == 4
    assert small["point_list"][1] == 0 and large["point_list"][1] == 0       # the sorted id list stays at offset 0
    assert large["ranges"][3] == 66049
    need = max(d["tile_sorted"][1] + d["tile_sorted"][2] * 5000 for d in (small, large))
    assert lib.ex4dgs_binning_bytes(5000) >= need
