"""CPU: the exact-output claim of the tile culling, checked on its mathematics.

oracle/cull_check.c restates tile_rect / tight_rect / cull_prepare / cull_test of ex4dgs_b200/csrc/common.cuh operation
by operation and verifies them by brute force: for every (splat, tile) instance of the reference's rectangle that the
bounding-box cut or the exact tile test drops, the compositing loop's own per-pixel arithmetic is evaluated at all 256
pixel centres of the tile (and around the +-pad subpixel box); a dropped instance with a contributing pixel would be a
violation.  The splats are built to hurt: needle-like conics at every angle, opacities spread around the 1/255
visibility limit, footprints from sub-pixel to hundreds of pixels, centres inside and far outside the image.
(The CUDA kernels themselves are compared with the exact-list mode on such scenes in tests/test_gpu_parity.py.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(os.path.dirname(HERE), "oracle")


def _lib():
    so = os.path.join(ORACLE, "libcullcheck.so")
    src = os.path.join(ORACLE, "cull_check.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE, "-B", "libcullcheck.so"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(so)
    lib.cull_check.restype = C.c_long
    lib.cull_check.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def _splats(n, seed, W=1352, H=1014):
    """Screen-space splats the way preprocess produces them (forward.cu:74-124, 222-250): filtered 2-D covariance ->
    conic (float32), radius = ceil(3 sqrt(lambda_max))."""
    g = np.random.default_rng(seed)
    s1 = np.exp(g.uniform(np.log(0.3), np.log(40.0), n))
    ratio = np.exp(g.uniform(0.0, np.log(300.0), n) * (g.random(n) < 0.6))       # 60 % needles, up to 300:1
    s2 = s1 / ratio
    big = g.random(n) < 0.02
    s1[big] *= 8.0                                                                # a few footprints of hundreds of pixels
    th = g.uniform(0.0, np.pi, n)
    c, s = np.cos(th), np.sin(th)
    a = (c * c * s1 * s1 + s * s * s2 * s2 + 0.1).astype(np.float32)
    b = (c * s * (s1 * s1 - s2 * s2)).astype(np.float32)
    cc = (s * s * s1 * s1 + c * c * s2 * s2 + 0.1).astype(np.float32)
    det = (a * cc - b * b).astype(np.float32)
    det_inv = (np.float32(1.0) / det).astype(np.float32)
    A, B, Cc = (cc * det_inv).astype(np.float32), (-b * det_inv).astype(np.float32), (a * det_inv).astype(np.float32)
    mid = np.float32(0.5) * (a + cc)
    lam = mid + np.sqrt(np.maximum(np.float32(0.1), mid * mid - det)).astype(np.float32)
    radius = np.ceil(np.float32(3.0) * np.sqrt(lam)).astype(np.int32)
    opac = np.where(g.random(n) < 0.4, (1.0 / 255.0) * np.exp(g.uniform(-0.7, 1.5, n)), g.uniform(0.0, 1.0, n) ** 2 + 1e-4)
    opac = np.minimum(opac, 1.0).astype(np.float32)
    cx = g.uniform(-300.0, W + 300.0, n).astype(np.float32)
    cy = g.uniform(-300.0, H + 300.0, n).astype(np.float32)
    return cx, cy, A, B, Cc, opac, radius


@pytest.mark.parametrize("pad,seed", [(0.0, 1), (0.0, 2), (0.5, 3), (3.0, 4)])
def test_dropped_instances_never_contribute(pad, seed):
    lib = _lib()
    n = 12000
    cx, cy, A, B, Cc, opac, radius = _splats(n, seed)
    counts = np.zeros(5, np.int64)
    first = C.c_int(-1)
    arrs = [np.ascontiguousarray(x) for x in (cx, cy, A, B, Cc, opac, radius)]
    bad = lib.cull_check(n, *[x.ctypes.data for x in arrs], pad, 85, 64, counts.ctypes.data, C.byref(first))
    ref_rect, tight, kept, need, viol = [int(v) for v in counts]
    print("pad %.1f: reference rectangles %d, bounding-box rectangles %d, kept by the exact test %d, contributing %d"
          % (pad, ref_rect, tight, kept, need))
    assert bad == 0 and viol == 0, "splat %d: a dropped (splat, tile) instance has a contributing pixel" % first.value
    assert ref_rect > 150000 and need > 10000                         # the sample exercises the test
    assert ref_rect >= tight >= kept >= need
    # ... and the culling is tight, not just safe: what it keeps beyond the contributing instances are sub-pixel
    # needles that cross a tile between its pixel centres (the test works on the tile's continuous rectangle)
    assert kept <= 2.0 * need


@pytest.mark.parametrize("bh,ox,oy,seed", [(4, 0.0, 0.0, 5), (8, 0.0, 0.0, 6), (4, 0.37, -0.45, 7), (8, -2.5, 3.0, 8)])
def test_block_reject_and_skip_threshold_never_drop_a_contribution(bh, ox, oy, seed):
    """The per-warp block test (8x4 pixel blocks in the forward, 8x8 in the backward) and the per-pair skip threshold
    of the compositing kernels against the compositing loop's own alpha test at every pixel."""
    lib = _lib()
    lib.block_check.restype = C.c_long
    lib.block_check.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]
    n = 5000
    arrs = [np.ascontiguousarray(x) for x in _splats(n, seed)]
    counts = np.zeros(6, np.int64)
    bad = lib.block_check(n, *[x.ctypes.data for x in arrs], bh, ox, oy, 85, 64, counts.ctypes.data)
    blocks, rejected, skipped, contributing, v_block, v_skip = [int(v) for v in counts]
    print("8x%d blocks %d, rejected %d (%.0f %%); pairs below the skip threshold %d, contributing %d"
          % (bh, blocks, rejected, 100.0 * rejected / blocks, skipped, contributing))
    assert bad == 0 and v_block == 0 and v_skip == 0
    assert rejected > 0.3 * blocks and contributing > 100000 and skipped > contributing
