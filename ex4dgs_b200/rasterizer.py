"""Python surface of the rasterizer: the reference's L1 API, bound to the C ABI with ctypes.

Mirrors submodules/diff_gaussian_rasterization_df/diff_gaussian_rasterization_df/__init__.py
(reference file:line in each docstring) - same class names, argument names and order, output
tuple order, gradient slots, empty-tensor conventions and exceptions - so that
gaussian_renderer/__init__.py:15,41-109 runs unchanged against it.  PyTorch is used only for
device memory, streams and autograd plumbing; all compute is in libex4dgs_raster.so.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib

# Process-wide default for the exact-output tile culling (include/ex4dgs_raster.h,
# EX4DGS_FLAG_TILE_CULL).  Outputs and gradients are identical either way; with 0 the internal
# tile lists are bit-identical to the reference's.  Default on; override with EX4DGS_TILE_CULL=0/1.
_DEFAULT_FLAGS = int(os.environ.get("EX4DGS_TILE_CULL", "1")) & 1


# EX4DGS_MATERIALIZE_GRADS=1: A/B knob - let autograd hand zero tensors to the backward for unused outputs
# (the reference's behaviour) instead of None.
_MATERIALIZE_GRADS = os.environ.get("EX4DGS_MATERIALIZE_GRADS", "0") == "1"


def set_default_flags(tile_cull: bool) -> None:
    global _DEFAULT_FLAGS
    _DEFAULT_FLAGS = _lib.FLAG_TILE_CULL if tile_cull else 0


def get_default_flags() -> int:
    return _DEFAULT_FLAGS


# Forward without any host wait (EX4DGS_FLAG_NO_HOST_WAIT): CUDA-graph capturable.  The binning buffer is sized from the
# capacity hint (set_capacity_hint, or the calling thread's earlier frames), `num_rendered` is that capacity, and a frame
# with more instances than the capacity is truncated and flagged on the device: check frame_overflowed() when the host
# synchronises anyway.  Off by default (the default forward waits for the instance count while the frame is queued and
# is always exact).
_HOST_WAIT = True


def set_host_wait(wait: bool) -> None:
    global _HOST_WAIT
    _HOST_WAIT = bool(wait)


def get_host_wait() -> bool:
    return _HOST_WAIT


def set_capacity_hint(instances: int) -> None:
    """Expected number of (Gaussian, tile) instances of the calling thread's next frames (0: forget the history)."""
    _lib.load().ex4dgs_set_capacity_hint(int(instances))


def frame_overflowed(out: torch.Tensor) -> bool:
    """Did the frame that produced `out` (an output of the rasterizer that requires grad) have more instances than its
    binning buffer held?  Only possible with set_host_wait(False).  Waits for the device."""
    word = getattr(out.grad_fn, "overflow_word", None)
    return bool(int(word.item()) & 2) if word is not None else False


def last_inexact_thresholds() -> int:
    """Diagnostics of the calling thread's last forward: visible Gaussians whose exact alpha >= 1/255 threshold could not
    be established (csrc/preprocess.cu alpha_threshold); 0 on everything observed so far."""
    return int(_lib.load().ex4dgs_last_inexact_thresholds())


def cpu_deep_copy_tuple(input_tuple):
    """__init__.py:18-20"""
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


def _ptr(t):
    """Device pointer of a tensor, NULL for the reference's 'absent' 0-element tensors
    (rasterize_points.cu passes data_ptr of a CPU empty tensor == nullptr)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t: torch.Tensor, dev) -> torch.Tensor:
    if t.numel() == 0:
        return t
    if t.device != dev:
        raise RuntimeError("ex4dgs_b200: tensor on %s, expected %s" % (t.device, dev))
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# Allocator callbacks handing out torch byte tensors (the C form of rasterize_points.cu:27-33
# resizeFunctional).  One persistent ctypes trampoline for the whole process: the `user` cookie is
# the buffer index (0 geometry, 1 binning, 2 image) and the tensors land in a per-thread list that
# the calling forward() installs - no per-call CFUNCTYPE objects and no reference cycles that
# would keep hundreds of MB of scratch alive until the cyclic GC runs.
_tls = threading.local()


def _alloc_trampoline(user, nbytes):
    i = int(user or 0)
    t = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=_tls.device)
    _tls.tensors[i] = t
    return t.data_ptr()


_ALLOC_CB = _lib.ALLOC_FN(_alloc_trampoline)


class SegmentedSH:
    """The model's SH coefficients as it stores them - features_dc [Ns,1,3] / features_rest [Ns,15,3] of the
    static Gaussians and the same pair of the dynamic ones - handed to the rasterizer WITHOUT the two
    levels of torch.cat of CGaussianModel.get_features() (scene/c_gaussian_model.py:337-353).  Pass it as
    `shs=`; the reference's render() only forwards `pc.get_features()` to the rasterizer
    (gaussian_renderer/__init__.py:95-103), so a model wrapper may return this object from get_features()
    (ex4dgs_b200.frontend.FusedGetters does).  Gradients flow to the four tensors directly."""

    def __init__(self, dc_static, rest_static, dc_dynamic, rest_dynamic):
        self.parts = (dc_static, rest_static, dc_dynamic, rest_dynamic)
        for t, k in zip(self.parts, (1, 15, 1, 15)):
            if t.dim() != 3 or t.shape[1] != k or t.shape[2] != 3:
                raise ValueError("SegmentedSH needs [N,1,3] / [N,15,3] tensors (SH degree 3 layout), got %s" % (tuple(t.shape),))
        if dc_static.shape[0] != rest_static.shape[0] or dc_dynamic.shape[0] != rest_dynamic.shape[0]:
            raise ValueError("SegmentedSH: dc / rest row counts differ")

    @property
    def shape(self):
        return (self.parts[0].shape[0] + self.parts[2].shape[0], 16, 3)

    def cat(self) -> torch.Tensor:
        """What get_features() would have returned."""
        dc_s, rest_s, dc_d, rest_d = self.parts
        return torch.cat((torch.cat((dc_s, rest_s), dim=1), torch.cat((dc_d, rest_d), dim=1)), dim=0)


def rasterize_gaussians(means3D, means2D, dir3D, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings):
    """__init__.py:22-45"""
    if isinstance(sh, SegmentedSH):
        return _RasterizeGaussiansSegmentedSH.apply(means3D, means2D, dir3D, *sh.parts, opacities, scales,
                                                    rotations, cov3Ds_precomp, raster_settings)
    return _RasterizeGaussians.apply(means3D, means2D, dir3D, sh, colors_precomp, opacities, scales,
                                     rotations, cov3Ds_precomp, raster_settings)


def _segments_struct(parts, dev):
    """(ctypes struct, tensors kept alive) for EX4DGS_FLAG_SH_SEGMENTED."""
    keep = []
    for t in parts:
        if t.numel() and not t.is_cuda:
            raise RuntimeError("ex4dgs_b200: SegmentedSH tensors must be CUDA tensors")
        keep.append(_f32c(t, dev))
    st = _lib.ShSegments()
    st.n_static = int(keep[0].shape[0])
    st.dc_static, st.rest_static, st.dc_dynamic, st.rest_dynamic = [_ptr(t) for t in keep]
    return st, keep


class _RasterizeGaussians(torch.autograd.Function):
    """__init__.py:47-178.  forward returns (color, radii, depth, flow, acc, idxs); backward returns
    the 10 slots (means3D, means2D, dir3D, sh, colors_precomp, opacities, scales, rotations,
    cov3Ds_precomp, None)."""

    @staticmethod
    def forward(ctx, means3D, means2D, dir3D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        return _forward_impl(ctx, means3D, means2D, dir3D, sh, colors_precomp, opacities, scales, rotations,
                             cov3Ds_precomp, raster_settings)

    @staticmethod
    def backward(ctx, grad_out_color, _, grad_out_depth, grad_out_flow, grad_out_acc, grad_out_idx):
        g = _backward_impl(ctx, grad_out_color, grad_out_depth, grad_out_flow, grad_out_acc)
        # slots whose forward input was "absent" get None (autograd ignores them in the reference too)
        return (g["means3D"], g["means2D"], g["dir3D"], g["sh"], g["colors"], g["opacities"], g["scales"], g["rotations"],
                g["cov3D"], None)


class _RasterizeGaussiansSegmentedSH(torch.autograd.Function):
    """Same operator with the SH input given as the model's four tensors (SegmentedSH): 12 slots
    (means3D, means2D, dir3D, dc_static, rest_static, dc_dynamic, rest_dynamic, opacities, scales, rotations,
    cov3Ds_precomp, None)."""

    @staticmethod
    def forward(ctx, means3D, means2D, dir3D, dc_s, rest_s, dc_d, rest_d, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        return _forward_impl(ctx, means3D, means2D, dir3D, (dc_s, rest_s, dc_d, rest_d), torch.Tensor([]), opacities,
                             scales, rotations, cov3Ds_precomp, raster_settings)

    @staticmethod
    def backward(ctx, grad_out_color, _, grad_out_depth, grad_out_flow, grad_out_acc, grad_out_idx):
        g = _backward_impl(ctx, grad_out_color, grad_out_depth, grad_out_flow, grad_out_acc)
        return (g["means3D"], g["means2D"], g["dir3D"], *g["sh_parts"], g["opacities"], g["scales"], g["rotations"],
                g["cov3D"], None)


def _forward_impl(ctx, means3D, means2D, dir3D, sh, colors_precomp, opacities, scales, rotations,
                  cov3Ds_precomp, raster_settings):
    rs = raster_settings
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")   # rasterize_points.cu:62-64
    if not means3D.is_cuda:
        raise RuntimeError("ex4dgs_b200: the rasterizer is CUDA-only (as is the reference, "
                           "rasterize_points.cu:80); got a %s tensor" % means3D.device)
    lib = _lib.load()
    dev = means3D.device
    flags = int(getattr(rs, "_flags", _DEFAULT_FLAGS))
    if not _HOST_WAIT:
        flags |= _lib.FLAG_NO_HOST_WAIT

    args = (rs.bg, means3D, dir3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
            cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.kernel_size,
            rs.subpixel_offset, rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos,
            rs.prefiltered, rs.min_depth, rs.max_depth, rs.debug)
    cpu_args = cpu_deep_copy_tuple(args) if rs.debug else None     # __init__.py:91-93

    P = means3D.size(0)
    H, W = int(rs.image_height), int(rs.image_width)
    means3D_c = _f32c(means3D, dev)
    dir3D_c = _f32c(dir3D, dev)
    seg, seg_keep = None, []
    if isinstance(sh, tuple):
        seg, seg_keep = _segments_struct(sh, dev)
        if seg_keep[0].shape[0] + seg_keep[2].shape[0] != means3D.size(0):
            raise RuntimeError("SegmentedSH holds %d Gaussians, means3D %d" % (seg_keep[0].shape[0] + seg_keep[2].shape[0], means3D.size(0)))
        flags |= _lib.FLAG_SH_SEGMENTED
        sh_c = torch.empty(0, dtype=torch.float32, device=dev)
    else:
        sh_c = _f32c(sh, dev)
    colors_c = _f32c(colors_precomp, dev)
    opac_c = _f32c(opacities, dev)
    scales_c = _f32c(scales, dev)
    rot_c = _f32c(rotations, dev)
    cov_c = _f32c(cov3Ds_precomp, dev)
    bg = _f32c(rs.bg, dev)
    view = _f32c(rs.viewmatrix, dev)
    proj = _f32c(rs.projmatrix, dev)
    campos = _f32c(rs.campos, dev)
    sub = _f32c(rs.subpixel_offset, dev)
    M = 16 if seg is not None else (sh_c.size(1) if sh_c.numel() != 0 else 0)        # rasterize_points.cu:92-96
    sh_arg = C.cast(C.pointer(seg), C.c_void_p) if seg is not None else _ptr(sh_c)

    iopt = dict(dtype=torch.int32, device=dev)
    fopt = dict(dtype=torch.float32, device=dev)
    if P == 0:
        # rasterize_points.cu:73-90: nothing runs, outputs keep their fill values
        color = torch.zeros(3, H, W, **fopt)
        radii = torch.zeros(0, **iopt)
        depth = torch.zeros(1, H, W, **fopt)
        acc = torch.zeros(1, H, W, **fopt)
        flow = torch.zeros(3, H, W, **fopt)
        idxs = torch.full((1, H, W), -1, **iopt)
        empty = torch.empty(0, dtype=torch.uint8, device=dev)
        ctx.raster_settings = rs
        ctx.num_rendered = 0
        ctx.flags = flags
        ctx.segmented = seg is not None
        ctx.save_for_backward(colors_c, means3D_c, scales_c, rot_c, cov_c, radii, sh_c, empty, empty, empty, depth, acc,
                              *seg_keep)
        return color, radii, depth, flow, acc, idxs

    color = torch.empty(3, H, W, **fopt)
    radii = torch.empty(P, **iopt)
    depth = torch.empty(1, H, W, **fopt)
    acc = torch.empty(1, H, W, **fopt)
    flow = torch.empty(3, H, W, **fopt)
    idxs = torch.empty(1, H, W, **iopt)
    _tls.device = dev
    _tls.tensors = [None, None, None]
    stream = torch.cuda.current_stream(dev).cuda_stream
    try:
        with torch.cuda.device(dev):
            R = lib.ex4dgs_forward(
                _ALLOC_CB, C.c_void_p(0), _ALLOC_CB, C.c_void_p(1), _ALLOC_CB, C.c_void_p(2),
                P, int(rs.sh_degree), int(M),
                _ptr(bg), W, H,
                _ptr(means3D_c), _ptr(dir3D_c), sh_arg, _ptr(colors_c),
                _ptr(opac_c), _ptr(scales_c), float(rs.scale_modifier), _ptr(rot_c),
                _ptr(cov_c), _ptr(view), _ptr(proj), _ptr(campos),
                float(rs.tanfovx), float(rs.tanfovy), float(rs.kernel_size), _ptr(sub), int(bool(rs.prefiltered)),
                _ptr(color), float(rs.min_depth), float(rs.max_depth), _ptr(depth), _ptr(acc), _ptr(flow),
                _ptr(idxs), _ptr(radii), int(bool(rs.debug)), flags, C.c_void_p(stream))
        if R < 0:
            raise RuntimeError("ex4dgs_forward failed (%d): %s" % (R, _lib.last_error()))
    except Exception as ex:
        if rs.debug:                                             # __init__.py:94-99
            torch.save(cpu_args, "snapshot_fw.dump")
            print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
        raise ex

    geomBuffer, binningBuffer, imgBuffer = _tls.tensors
    _tls.tensors = None
    # outputs the loss does not use arrive as None in backward (instead of the zero tensors the reference
    # receives, __init__.py:110): the C ABI takes NULL for them and skips their terms
    if not _MATERIALIZE_GRADS:
        ctx.set_materialize_grads(False)
    ctx.raster_settings = rs
    ctx.num_rendered = int(R)
    if flags & _lib.FLAG_NO_HOST_WAIT:
        # device word meta[5] of the geometry buffer (bit 1: the tile lists were truncated), see frame_overflowed()
        _, off, _, _ = _lib.describe_buffers(P, 0, W, H)["meta"]
        a = ((geomBuffer.data_ptr() + 255) & ~255) - geomBuffer.data_ptr()
        ctx.overflow_word = geomBuffer[a + off + 20:a + off + 24].view(torch.int32)
    ctx.flags = flags
    ctx.segmented = seg is not None
    ctx.save_for_backward(colors_c, means3D_c, scales_c, rot_c, cov_c, radii, sh_c,
                          geomBuffer, binningBuffer, imgBuffer, depth, acc, *seg_keep)
    ctx.mark_non_differentiable(radii, idxs)
    return color, radii, depth, flow, acc, idxs


def _backward_impl(ctx, grad_out_color, grad_out_depth, grad_out_flow, grad_out_acc):
    rs = ctx.raster_settings
    (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh,
     geomBuffer, binningBuffer, imgBuffer, depth, acc) = ctx.saved_tensors[:12]
    seg_parts = ctx.saved_tensors[12:] if getattr(ctx, "segmented", False) else None
    lib = _lib.load()
    dev = means3D.device
    P = means3D.size(0)
    H, W = int(rs.image_height), int(rs.image_width)
    M = 16 if seg_parts is not None else (sh.size(1) if sh.numel() != 0 else 0)
    fopt = dict(dtype=torch.float32, device=dev)

    args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, depth, acc, rs.min_depth, rs.max_depth,
            rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
            rs.kernel_size, rs.subpixel_offset, grad_out_color, grad_out_depth, grad_out_flow, grad_out_acc,
            sh, rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, rs.debug)
    cpu_args = cpu_deep_copy_tuple(args) if rs.debug else None   # __init__.py:151-153

    use_sr = scales.numel() != 0
    g_means2D = torch.empty(P, 3, **fopt)
    g_colors = torch.empty(P, 3, **fopt)
    g_opac = torch.empty(P, 1, **fopt)
    g_means3D = torch.empty(P, 3, **fopt)
    g_cov3D = torch.empty(P, 6, **fopt) if not use_sr else torch.empty(0, **fopt)
    g_sh = torch.empty(P, M, 3, **fopt) if seg_parts is None else None
    g_parts = [torch.empty_like(t) for t in seg_parts] if seg_parts is not None else None
    sh_in, sh_out = _ptr(sh), _ptr(g_sh)
    if seg_parts is not None:
        seg_in, _ = _segments_struct(seg_parts, dev)
        seg_out, _ = _segments_struct(g_parts, dev)
        sh_in = C.cast(C.pointer(seg_in), C.c_void_p)
        sh_out = C.cast(C.pointer(seg_out), C.c_void_p)
    g_scales = torch.empty(P, 3, **fopt) if use_sr else torch.zeros(P, 3, **fopt)
    g_rot = torch.empty(P, 4, **fopt) if use_sr else torch.zeros(P, 4, **fopt)
    g_dir = torch.empty(P, 3, **fopt)

    if P != 0:
        gc = _f32c(grad_out_color, dev) if grad_out_color is not None else torch.zeros(3, H, W, **fopt)
        gd = _f32c(grad_out_depth, dev) if grad_out_depth is not None else None
        gf = _f32c(grad_out_flow, dev) if grad_out_flow is not None else None
        ga = _f32c(grad_out_acc, dev) if grad_out_acc is not None else None
        bg = _f32c(rs.bg, dev)
        view = _f32c(rs.viewmatrix, dev)
        proj = _f32c(rs.projmatrix, dev)
        campos = _f32c(rs.campos, dev)
        sub = _f32c(rs.subpixel_offset, dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        try:
            with torch.cuda.device(dev):
                rc = lib.ex4dgs_backward(
                    P, int(rs.sh_degree), int(M), int(ctx.num_rendered),
                    _ptr(bg), W, H,
                    _ptr(means3D), sh_in, _ptr(colors_precomp),
                    _ptr(scales), float(rs.scale_modifier), _ptr(rotations),
                    _ptr(depth), _ptr(acc), float(rs.min_depth), float(rs.max_depth),
                    _ptr(cov3Ds_precomp), _ptr(view), _ptr(proj), _ptr(campos),
                    float(rs.tanfovx), float(rs.tanfovy), float(rs.kernel_size), _ptr(sub), _ptr(radii),
                    _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer),
                    _ptr(gc), _ptr(gd), _ptr(gf), _ptr(ga),
                    _ptr(g_means2D), _ptr(g_opac), _ptr(g_colors), _ptr(g_means3D), _ptr(g_cov3D),
                    sh_out, _ptr(g_scales) if use_sr else None, _ptr(g_rot) if use_sr else None, _ptr(g_dir),
                    int(bool(rs.debug)), int(ctx.flags), C.c_void_p(stream))
            if rc < 0:
                raise RuntimeError("ex4dgs_backward failed (%d): %s" % (rc, _lib.last_error()))
        except Exception as ex:
            if rs.debug:                                         # __init__.py:156-159
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
            raise ex

    return dict(means3D=g_means3D, means2D=g_means2D, dir3D=g_dir,
                sh=g_sh if (M != 0 and seg_parts is None) else None, sh_parts=g_parts,
                colors=g_colors if colors_precomp.numel() != 0 else None, opacities=g_opac,
                scales=g_scales if use_sr else None, rotations=g_rot if use_sr else None,
                cov3D=g_cov3D if not use_sr else None)


class GaussianRasterizationSettings(NamedTuple):
    """__init__.py:180-196 (field order is part of the contract)."""
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


class GaussianRasterizer(nn.Module):
    """__init__.py:198-251."""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Frustum test per Gaussian (rasterizer_impl.cu:54-68).  The reference's Python wrapper
        passes 4 arguments to a 5-argument native function (__init__.py:207-211 vs
        rasterize_points.h:78-83) and therefore raises TypeError; this one works and uses the
        settings' min_depth / max_depth."""
        with torch.no_grad():
            rs = self.raster_settings
            lib = _lib.load()
            dev = positions.device
            if not positions.is_cuda:
                raise RuntimeError("ex4dgs_b200: markVisible is CUDA-only")
            pos = _f32c(positions, dev)
            P = pos.size(0)
            present = torch.zeros(P, dtype=torch.bool, device=dev)
            if P != 0:
                view = _f32c(rs.viewmatrix, dev)
                proj = _f32c(rs.projmatrix, dev)
                stream = torch.cuda.current_stream(dev).cuda_stream
                with torch.cuda.device(dev):
                    rc = lib.ex4dgs_mark_visible(P, _ptr(pos), _ptr(view), _ptr(proj), float(rs.min_depth),
                                                 float(rs.max_depth), present.data_ptr(), C.c_void_p(stream))
                if rc < 0:
                    raise RuntimeError("ex4dgs_mark_visible failed: " + _lib.last_error())
        return present

    def forward(self, means3D, means2D, dir3D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        if dir3D is None:
            dir3D = torch.Tensor([])

        return rasterize_gaussians(means3D, means2D, dir3D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, raster_settings)
