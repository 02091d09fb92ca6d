"""Per-iteration bookkeeping of the training loop ("stats updates").

`iteration_stats(model, ...)` replaces the block of train.py:196-215 that runs after `loss.backward()`:

    if opt.l1_accum: gaussians.mark_prune_stats(radii, viewspace_point_error_tensor)       # c_gaussian_model.py:1105
    if iteration < opt.densify_until_iter:
        gaussians.max_radii2D[static_vis] = torch.max(...)                                  # train.py:205-206
        gaussians.motion_max_radii2D[dynamic_vis] = torch.max(...)
        gaussians.add_densification_stats(viewspace_point_tensor, ...)                      # c_gaussian_model.py:1095
        if opt.l1_accum: gaussians.add_l1_ssim_stats(viewspace_point_error_tensor, ...)     # c_gaussian_model.py:1119

on the model's own statistics tensors (same attribute names, updated in place), in one CUDA launch and
without the host synchronisation every boolean-mask index of the reference implies.

`regularizers_(...)` evaluates the default-on regularisation terms of the loss (train.py:156-162) and adds
their gradients to the `.grad` the backward pass has produced.

CUDA only; no fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

# (field of ex4dgs_stats_arrays, attribute of the static model part, attribute of the dynamic part)
_STAT_FIELDS = (
    ("max_radii2D", "max_radii2D", "motion_max_radii2D"),
    ("min_radii2D", "min_radii2D", "motion_min_radii2D"),
    ("xyz_gradient_accum", "xyz_gradient_accum", "motion_xyz_gradient_accum"),
    ("denom", "denom", "motion_denom"),
    ("error_accum", "xyz_error_accum", "motion_xyz_error_mean"),
    ("error_min", "xyz_error_min", "motion_xyz_error_min"),
    ("error_min_timestamp", "xyz_error_min_timestamp", "motion_xyz_error_min_timestamp"),
    ("ssim_error_accum", "xyz_ssim_error_accum", "motion_xyz_ssim_error_accum"),
    ("error_denom", "error_denom", "motion_error_denom"),
)


def _check(t: torch.Tensor, n: int, name: str, dev) -> torch.Tensor:
    if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n:
        raise RuntimeError("ex4dgs_b200.stats: %s must be a contiguous float32 tensor with %d elements on %s "
                           "(got %s %s on %s)" % (name, n, dev, tuple(t.shape), t.dtype, t.device))
    return t


def _arrays(model, which: int, n: int, dev, need):
    st = _lib.StatsArrays()
    for field, *attrs in _STAT_FIELDS:
        attr = attrs[which]
        t = getattr(model, attr, None)
        if t is None or n == 0:
            if n != 0 and field in need:
                raise AttributeError("ex4dgs_b200.stats: the model has no `%s`" % attr)
            setattr(st, field, None)
        else:
            setattr(st, field, _check(t, n, attr, dev).data_ptr())
    return st


def iteration_stats(model, radii: torch.Tensor, viewspace_point_tensor_grad: torch.Tensor,
                    viewspace_point_error_tensor_grad: Optional[torch.Tensor], timestamp: float,
                    densify: bool = True, static_num: Optional[int] = None) -> None:
    """train.py:196-215 in one launch.  `model` carries the statistics tensors under the reference's attribute
    names (CGaussianModel.training_setup, scene/c_gaussian_model.py:412-428, and :408-409/:843-844 for the
    radii); `viewspace_point_tensor_grad` is `render_pkg["viewspace_points"].grad`, the error gradient is
    `render_pkg["viewspace_l1points"].grad` or None when opt.l1_accum is off; `densify` is
    `iteration < opt.densify_until_iter`."""
    if not radii.is_cuda:
        raise RuntimeError("ex4dgs_b200: iteration_stats is CUDA-only")
    dev = radii.device
    P = radii.numel()
    Ns = int(static_num if static_num is not None else model._xyz.shape[0])
    Nd = P - Ns
    if Nd < 0:
        raise RuntimeError("iteration_stats: %d radii for %d static Gaussians" % (P, Ns))
    if radii.dtype != torch.int32 or not radii.is_contiguous():
        raise RuntimeError("iteration_stats: radii must be the rasterizer's contiguous int32 output")
    g = _check(viewspace_point_tensor_grad, 3 * P, "viewspace_point_tensor.grad", dev)
    e = viewspace_point_error_tensor_grad
    if e is not None:
        e = _check(e, 3 * P, "viewspace_point_error_tensor.grad", dev)
    need = set()
    if e is not None:
        need.add("min_radii2D")
    if densify:
        need |= {"max_radii2D", "xyz_gradient_accum", "denom"}
        if e is not None:
            need |= {"error_accum", "error_min", "error_min_timestamp", "ssim_error_accum", "error_denom"}
    stat = _arrays(model, 0, Ns, dev, need)
    dyn = _arrays(model, 1, Nd, dev, need)
    lib = _lib.load()
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = lib.ex4dgs_iteration_stats(Ns, Nd, radii.data_ptr(), g.data_ptr(), e.data_ptr() if e is not None else None,
                                        float(timestamp), int(bool(densify)), C.byref(stat), C.byref(dyn), C.c_void_p(stream))
    if rc < 0:
        raise RuntimeError("ex4dgs_iteration_stats failed (%d): %s" % (rc, _lib.last_error()))


_scratch = {}


def regularizers_(xyz_disp: torch.Tensor, xyz_motion: torch.Tensor, static_reg: float, motion_reg: float,
                  loss_grad: Optional[torch.Tensor] = None, value_only: bool = False) -> torch.Tensor:
    """The regularisation terms train.py:156-162 adds to the loss:

        static_reg * torch.log(gaussians._xyz_disp.norm(dim=-1) + 0.001).mean()
        motion_reg * (gaussians._xyz_motion[:, :1] - gaussians._xyz_motion[:, 1:]).norm(dim=-1).mean()

    Returns a float32 tensor [2] with the two terms (0 where the weight is 0 or the tensor is empty) and, unless
    `value_only`, ADDS their gradients (times `loss_grad`, a device scalar, default 1) to `xyz_disp.grad` /
    `xyz_motion.grad` in place (allocating them when they are None) - call it after `loss.backward()`; the sum of
    the returned terms is what the reference adds to `loss`."""
    if not xyz_disp.is_cuda:
        raise RuntimeError("ex4dgs_b200: regularizers_ is CUDA-only")
    dev = xyz_disp.device
    Ns = int(xyz_disp.shape[0])
    Nd = int(xyz_motion.shape[0])
    K = int(xyz_motion.shape[1]) if xyz_motion.dim() == 3 else 0
    for t, name in ((xyz_disp, "xyz_disp"), (xyz_motion, "xyz_motion")):
        if t.numel() and (t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev):
            raise RuntimeError("regularizers_: %s must be a contiguous float32 tensor on %s" % (name, dev))
    lib = _lib.load()
    out = torch.empty(2, dtype=torch.float32, device=dev)
    sc = _scratch.get(dev)
    if sc is None:
        sc = _scratch[dev] = torch.empty(int(lib.ex4dgs_regularizer_scratch_bytes()), dtype=torch.uint8, device=dev)
    gd = gm = None
    acc_d = acc_m = 0
    if not value_only:
        if static_reg != 0 and Ns:
            acc_d = int(xyz_disp.grad is not None)
            if xyz_disp.grad is None:
                xyz_disp.grad = torch.empty_like(xyz_disp)
            gd = xyz_disp.grad
        if motion_reg != 0 and Nd and K > 1:
            acc_m = int(xyz_motion.grad is not None)
            if xyz_motion.grad is None:
                xyz_motion.grad = torch.empty_like(xyz_motion)
            gm = xyz_motion.grad
        for g in (gd, gm):
            if g is not None and (g.dtype != torch.float32 or not g.is_contiguous()):
                raise RuntimeError("regularizers_: gradients must be contiguous float32")
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = lib.ex4dgs_regularizers(Ns, Nd, K, xyz_disp.data_ptr() if Ns else None, xyz_motion.data_ptr() if Nd else None,
                                     float(static_reg), float(motion_reg),
                                     loss_grad.data_ptr() if loss_grad is not None else None,
                                     gd.data_ptr() if gd is not None else None, acc_d,
                                     gm.data_ptr() if gm is not None else None, acc_m,
                                     out.data_ptr(), sc.data_ptr(), C.c_void_p(stream))
    if rc < 0:
        raise RuntimeError("ex4dgs_regularizers failed (%d): %s" % (rc, _lib.last_error()))
    return out
