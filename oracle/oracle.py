"""ctypes wrapper around oracle/liboracle.so (cpu_raster.c) - TEST INFRASTRUCTURE.

Exposes the CPU restatement of the reference rasterizer
  * as plain functions on numpy arrays (`forward`, `backward`, `state`), and
  * behind the reference's L1 API (GaussianRasterizationSettings / GaussianRasterizer, CPU tensors,
    torch.autograd.Function with the hand-written backward) so that the parity tests read like
    tests of the reference and so that config 1 of BASELINE.json ("PyTorch-CPU autograd raster",
    which does not exist in the reference: rasterize_points.cu:80 is CUDA-only) can be timed.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, NamedTuple, Optional

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liboracle.so")

_lib = None
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "cpu_raster.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        _lib.or_create.restype = C.c_void_p
        _lib.or_destroy.argtypes = [C.c_void_p]
        _lib.or_forward.restype = C.c_int
        _lib.or_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int] + \
            [C.c_void_p] * 6 + [C.c_float, C.c_void_p, C.c_void_p] + [C.c_void_p] * 3 + [C.c_float] * 3 + \
            [C.c_void_p, C.c_float, C.c_float] + [C.c_void_p] * 6
        _lib.or_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 13
        _lib.or_backward.restype = C.c_int
        _lib.or_backward.argtypes = [C.c_void_p] + [C.c_void_p] * 14
        _lib.or_mark_visible.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    return _lib


def _f(a) -> Optional[np.ndarray]:
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if a.size else None


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One frame: forward() then optionally state() / backward()."""

    def __init__(self):
        self.h = lib().or_create()
        self._keep = []

    def __del__(self):
        try:
            lib().or_destroy(self.h)
        except Exception:
            pass

    def forward(self, *, bg, W, H, means3D, dir3D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None, viewmatrix, projmatrix, campos, tanfovx, tanfovy,
                kernel_size, subpixel_offset, min_depth, max_depth, sh_degree=3, scale_modifier=1.0) -> Dict[str, np.ndarray]:
        means3D = _f(means3D)
        P = 0 if means3D is None else means3D.shape[0]
        shs_ = _f(shs)
        M = shs_.shape[1] if shs_ is not None else 0
        arrs = dict(bg=_f(bg), means=means3D, dir=_f(dir3D), shs=shs_, col=_f(colors_precomp), opac=_f(opacities),
                    scales=_f(scales), rots=_f(rotations), cov=_f(cov3D_precomp), view=_f(viewmatrix), proj=_f(projmatrix),
                    cam=_f(campos), sub=_f(subpixel_offset))
        if arrs["sub"] is None:
            arrs["sub"] = np.zeros((H, W, 2), np.float32)
        self._keep = arrs      # the C context borrows these pointers until the next forward
        self.P, self.W, self.H, self.M = P, W, H, M
        out = dict(color=np.empty((3, H, W), np.float32), depth=np.empty((1, H, W), np.float32),
                   acc=np.empty((1, H, W), np.float32), flow=np.empty((3, H, W), np.float32),
                   idxs=np.empty((1, H, W), np.int32), radii=np.zeros(P, np.int32))
        R = lib().or_forward(self.h, P, int(sh_degree), M, _p(arrs["bg"]), W, H, _p(arrs["means"]), _p(arrs["dir"]),
                             _p(arrs["shs"]), _p(arrs["col"]), _p(arrs["opac"]), _p(arrs["scales"]), float(scale_modifier),
                             _p(arrs["rots"]), _p(arrs["cov"]), _p(arrs["view"]), _p(arrs["proj"]), _p(arrs["cam"]),
                             float(tanfovx), float(tanfovy), float(kernel_size), _p(arrs["sub"]), float(min_depth),
                             float(max_depth), _p(out["color"]), _p(out["depth"]), _p(out["acc"]), _p(out["flow"]),
                             _p(out["idxs"]), _p(out["radii"]))
        if R < 0:
            raise RuntimeError("oracle forward failed")
        self.R = R
        out["R"] = R
        return out

    def state(self) -> Dict[str, np.ndarray]:
        P, W, H, R = self.P, self.W, self.H, self.R
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        s = dict(depths=np.zeros(P, np.float32), means2D=np.zeros((P, 2), np.float32), cov3D=np.zeros((P, 6), np.float32),
                 conic_opacity=np.zeros((P, 4), np.float32), rgb=np.zeros((P, 3), np.float32), clamped=np.zeros((P, 3), np.uint8),
                 tiles_touched=np.zeros(P, np.uint32), point_offsets=np.zeros(P, np.uint32),
                 point_list_keys=np.zeros(R, np.uint64), point_list=np.zeros(R, np.uint32),
                 ranges=np.zeros((tiles, 2), np.uint32), final_T=np.zeros(W * H, np.float32), n_contrib=np.zeros(W * H, np.uint32))
        if P:
            lib().or_get_state(self.h, *[_p(s[k]) for k in ("depths", "means2D", "cov3D", "conic_opacity", "rgb", "clamped",
                                                            "tiles_touched", "point_offsets", "point_list_keys", "point_list",
                                                            "ranges", "final_T", "n_contrib")])
        s["R"] = R
        return s

    def backward(self, grad_color, grad_depth, grad_flow, grad_acc) -> Dict[str, np.ndarray]:
        P, M = self.P, self.M
        g = dict(means2D=np.zeros((P, 3), np.float32), colors=np.zeros((P, 3), np.float32), opacities=np.zeros((P, 1), np.float32),
                 means3D=np.zeros((P, 3), np.float32), cov3D=np.zeros((P, 6), np.float32), shs=np.zeros((P, M, 3), np.float32),
                 scales=np.zeros((P, 3), np.float32), rotations=np.zeros((P, 4), np.float32), dir3D=np.zeros((P, 3), np.float32),
                 conic=np.zeros((P, 4), np.float32))
        gi = [_f(grad_color), _f(grad_depth), _f(grad_flow), _f(grad_acc)]
        lib().or_backward(self.h, *[_p(a) for a in gi],
                          *[_p(g[k]) for k in ("means2D", "colors", "opacities", "means3D", "cov3D", "shs", "scales",
                                               "rotations", "dir3D", "conic")])
        return g


def mark_visible(means3D, viewmatrix, projmatrix, min_depth, max_depth) -> np.ndarray:
    m = _f(means3D)
    P = 0 if m is None else m.shape[0]
    out = np.zeros(P, np.uint8)
    if P:
        v, p = _f(viewmatrix), _f(projmatrix)
        lib().or_mark_visible(P, _p(m), _p(v), _p(p), float(min_depth), float(max_depth), _p(out))
    return out.astype(bool)


# ---- the reference's L1 API on CPU tensors (RAST/diff_gaussian_rasterization_df/__init__.py:22-251) ----

class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    kernel_size: float
    subpixel_offset: torch.Tensor
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    min_depth: float
    max_depth: float
    debug: bool


def _np(t):
    return None if t is None or t.numel() == 0 else t.detach().cpu().float().numpy()


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, dir3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
        o = Oracle()
        out = o.forward(bg=_np(rs.bg), W=int(rs.image_width), H=int(rs.image_height), means3D=_np(means3D), dir3D=_np(dir3D),
                        opacities=_np(opacities), shs=_np(sh), colors_precomp=_np(colors_precomp), scales=_np(scales),
                        rotations=_np(rotations), cov3D_precomp=_np(cov3Ds_precomp), viewmatrix=_np(rs.viewmatrix),
                        projmatrix=_np(rs.projmatrix), campos=_np(rs.campos), tanfovx=rs.tanfovx, tanfovy=rs.tanfovy,
                        kernel_size=rs.kernel_size, subpixel_offset=_np(rs.subpixel_offset), min_depth=rs.min_depth,
                        max_depth=rs.max_depth, sh_degree=rs.sh_degree, scale_modifier=rs.scale_modifier)
        ctx.oracle = o
        ctx.num_rendered = out["R"]
        ctx.flags_present = (sh.numel() != 0, colors_precomp.numel() != 0, scales.numel() != 0)
        t = {k: torch.from_numpy(out[k]) for k in ("color", "radii", "depth", "flow", "acc", "idxs")}
        ctx.mark_non_differentiable(t["radii"], t["idxs"])
        return t["color"], t["radii"], t["depth"], t["flow"], t["acc"], t["idxs"]

    @staticmethod
    def backward(ctx, g_color, _, g_depth, g_flow, g_acc, g_idx):
        g = ctx.oracle.backward(_np(g_color), _np(g_depth), _np(g_flow), _np(g_acc))
        has_sh, has_col, has_sr = ctx.flags_present
        T = torch.from_numpy
        return (T(g["means3D"]), T(g["means2D"]), T(g["dir3D"]), T(g["shs"]) if has_sh else None,
                T(g["colors"]) if has_col else None, T(g["opacities"]), T(g["scales"]) if has_sr else None,
                T(g["rotations"]) if has_sr else None, T(g["cov3D"]) if not has_sr else None, None)


class GaussianRasterizer(torch.nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        rs = self.raster_settings
        return torch.from_numpy(mark_visible(_np(positions), _np(rs.viewmatrix), _np(rs.projmatrix), rs.min_depth, rs.max_depth))

    def forward(self, means3D, means2D, dir3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        e = torch.Tensor([])
        return _RasterizeGaussians.apply(means3D, means2D, dir3D, e if shs is None else shs,
                                         e if colors_precomp is None else colors_precomp, opacities,
                                         e if scales is None else scales, e if rotations is None else rotations,
                                         e if cov3D_precomp is None else cov3D_precomp, self.raster_settings)
