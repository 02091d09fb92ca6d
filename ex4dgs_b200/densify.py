"""Tensor surgery of densification and pruning in one launch (csrc/compact.cu).

Every `densification_interval` iterations the reference rebuilds its 15 parameter tensors, the two RAdam moments of
each and its 18 per-Gaussian statistics tensors - by boolean mask (`CGaussianModel._prune_optimizer` /
`prune_points`, scene/c_gaussian_model.py:693-763) and by concatenation (`cat_tensors_to_optimizer`, :765-787, called
from `densification_postfix*`, :789-872, with the rows `densify_and_clone` / `densify_and_split` selected, :874-1017):
one `x[mask]` (a `nonzero` + host wait + gather) or one `torch.cat` per tensor, ~90 launches and ~45 host waits.

Here every one of these calls is ONE `ex4dgs_gather_rows` launch over a table of jobs (plus one `nonzero` per mask).
The three methods keep the reference's names, arguments, return values and side effects, so they can be bound onto a
reference model object and the rest of its densification code (`densify_and_prune`, `densify_and_clone`,
`densify_and_split`, `prune_invisible`, `prune_small`, `prune_nan_points`, `extract_dynamic_points_from_static`)
runs unchanged on top of them:

    ex4dgs_b200.densify.install(gaussians)        # after gaussians.training_setup(...)

They only touch `optimizer.param_groups` / `optimizer.state` (`step`, `exp_avg`, `exp_avg_sq`), which
`torch.optim.RAdam` and `FusedRAdam` share.  CUDA float32 tensors only - there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import types
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib

# the statistics tensors prune_points compacts (scene/c_gaussian_model.py:733-741, :755-763)
STATIC_STATS = ("xyz_gradient_accum", "denom", "xyz_error_accum", "xyz_error_min", "xyz_error_min_timestamp",
                "xyz_ssim_error_accum", "error_denom", "max_radii2D", "min_radii2D")
DYNAMIC_STATS = ("motion_xyz_gradient_accum", "motion_denom", "motion_xyz_error_min", "motion_xyz_error_mean",
                 "motion_xyz_error_min_timestamp", "motion_xyz_ssim_error_accum", "motion_error_denom",
                 "motion_max_radii2D", "motion_min_radii2D")
_STATIC_ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
                "scaling": "_scaling", "rotation": "_rotation", "xyz_disp": "_xyz_disp"}
_DYNAMIC_ATTR = {"motion_xyz": "_xyz_motion", "motion_f_dc": "_features_dc_motion", "motion_f_rest": "_features_rest_motion",
                 "motion_scaling": "_scaling_motion", "motion_opacity": "_opacity_motion",
                 "motion_opacity_center": "_opacity_duration_center", "motion_opacity_var": "_opacity_duration_var",
                 "motion_rotation": "_rotation_motion"}


class _Jobs:
    """A table of row-gather jobs: dst[r] = r < n_a ? a[index ? index[r] : r] : (b ? b[r - n_a] : 0)."""

    def __init__(self):
        self.jobs: List[tuple] = []
        self.keep: List[torch.Tensor] = []

    def add(self, a: torch.Tensor, n_out: int, index: Optional[torch.Tensor] = None, n_a: Optional[int] = None,
            b: Optional[torch.Tensor] = None, rows: Optional[int] = None) -> torch.Tensor:
        """Queue one job; returns the (still unwritten) output tensor of shape [n_out, *a.shape[1:]].
        rows: number of rows the index was built for (the mask's length) - the kernel does not range-check the index,
        so a source of another length is refused here, as torch refuses `x[mask]` with a mask of the wrong shape."""
        if not a.is_cuda:
            raise RuntimeError("ex4dgs_b200.densify is CUDA-only (no CPU fallback)")
        if rows is not None and (a.dim() == 0 or a.shape[0] != rows):
            raise IndexError("The shape of the mask [%d] at index 0 does not match the shape of the indexed tensor %s at index 0"
                             % (rows, list(a.shape)))
        src = a.detach()
        if not src.is_contiguous():
            src = src.contiguous()
        row = 1
        for s in src.shape[1:]:
            row *= int(s)
        row_bytes = row * src.element_size()
        n_a = n_out if n_a is None else int(n_a)
        out = torch.empty((n_out,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        if n_out == 0:
            return out
        if row_bytes == 0 or row_bytes % 4 != 0:
            raise RuntimeError("ex4dgs_b200.densify: rows must be a positive multiple of 4 bytes (got %d)" % row_bytes)
        ext = None
        if b is not None:
            ext = b.detach()
            if ext.dtype != src.dtype or tuple(ext.shape[1:]) != tuple(src.shape[1:]) or ext.device != src.device:
                raise RuntimeError("ex4dgs_b200.densify: extension %s %s does not match %s %s" %
                                   (tuple(ext.shape), ext.dtype, tuple(src.shape), src.dtype))
            if not ext.is_contiguous():
                ext = ext.contiguous()
            if ext.shape[0] != n_out - n_a:
                raise RuntimeError("ex4dgs_b200.densify: extension has %d rows, expected %d" % (ext.shape[0], n_out - n_a))
        if index is not None:
            if index.dtype != torch.int64 or not index.is_contiguous() or index.device != src.device or index.numel() < n_a:
                raise RuntimeError("ex4dgs_b200.densify: index must be a contiguous int64 tensor on the source's device")
        elif n_a > src.shape[0]:
            raise RuntimeError("ex4dgs_b200.densify: n_a=%d exceeds the %d source rows" % (n_a, src.shape[0]))
        self.jobs.append((src, ext, out, index, row_bytes, n_a, n_out))
        self.keep += [src, out] + ([ext] if ext is not None else []) + ([index] if index is not None else [])
        return out

    def launch(self) -> None:
        if not self.jobs:
            return
        lib = _lib.load()
        dev = self.jobs[0][0].device
        stream = torch.cuda.current_stream(dev).cuda_stream
        for i in range(0, len(self.jobs), 64):
            chunk = self.jobs[i:i + 64]
            arr = (_lib.GatherJob * len(chunk))()
            for d, (src, ext, out, index, row_bytes, n_a, n_out) in zip(arr, chunk):
                d.a = src.data_ptr() if src.numel() else None
                d.b = ext.data_ptr() if (ext is not None and ext.numel()) else None
                d.dst = out.data_ptr()
                d.index = index.data_ptr() if index is not None else None
                d.row_bytes, d.n_a, d.n_out = row_bytes, n_a, n_out
            with torch.cuda.device(dev):
                rc = lib.ex4dgs_gather_rows(arr, len(chunk), C.c_void_p(stream))
            if rc < 0:
                raise RuntimeError("ex4dgs_gather_rows failed (%d): %s" % (rc, _lib.last_error()))
        self.jobs, self.keep = [], []


def gather_rows(tensors: Sequence[torch.Tensor], index: torch.Tensor) -> List[torch.Tensor]:
    """[t[index] for t in tensors] (row gather along dim 0, int64 index) in one launch."""
    jobs = _Jobs()
    n = int(index.numel())
    outs = [jobs.add(t, n, index=index) for t in tensors]
    jobs.launch()
    return outs


def _keep_index(mask: torch.Tensor) -> torch.Tensor:
    return torch.nonzero(mask.reshape(-1)).reshape(-1)


def _rebuild_group(optimizer, group, new_param: torch.Tensor, new_avg, new_sq) -> nn.Parameter:
    """The bookkeeping of c_gaussian_model.py:701-712 / :774-785 for one parameter group."""
    old = group["params"][0]
    stored_state = optimizer.state.get(old, None)
    param = nn.Parameter(new_param.requires_grad_(True))
    if stored_state is not None:
        stored_state["exp_avg"] = new_avg
        stored_state["exp_avg_sq"] = new_sq
        del optimizer.state[old]
        group["params"][0] = param
        optimizer.state[param] = stored_state
    else:
        group["params"][0] = param
    return param


def _prune(optimizer, static_mask: torch.Tensor, dynamic_mask: torch.Tensor, extra_static=(), extra_dynamic=()):
    idx = {False: _keep_index(static_mask), True: _keep_index(dynamic_mask)}
    rows = {False: int(static_mask.numel()), True: int(dynamic_mask.numel())}
    jobs = _Jobs()
    planned = []
    for group in optimizer.param_groups:
        dyn = group["name"].startswith("motion_")
        index = idx[dyn]
        p = group["params"][0]
        n = int(index.numel())
        st = optimizer.state.get(p, None)
        new_p = jobs.add(p, n, index=index, rows=rows[dyn])
        new_avg = jobs.add(st["exp_avg"], n, index=index, rows=rows[dyn]) if st is not None else None
        new_sq = jobs.add(st["exp_avg_sq"], n, index=index, rows=rows[dyn]) if st is not None else None
        planned.append((group, new_p, new_avg, new_sq))
    outs_s = [jobs.add(t, int(idx[False].numel()), index=idx[False], rows=rows[False]) for t in extra_static]
    outs_d = [jobs.add(t, int(idx[True].numel()), index=idx[True], rows=rows[True]) for t in extra_dynamic]
    jobs.launch()
    optimizable_tensors = {}
    for group, new_p, new_avg, new_sq in planned:
        optimizable_tensors[group["name"]] = _rebuild_group(optimizer, group, new_p, new_avg, new_sq)
    return optimizable_tensors, outs_s, outs_d


def _prune_optimizer(self, static_mask, dynamic_mask) -> Dict[str, nn.Parameter]:
    """Drop-in for CGaussianModel._prune_optimizer (scene/c_gaussian_model.py:693-713): the masks select the rows to KEEP;
    groups whose name starts with "motion_" take the dynamic mask."""
    return _prune(self.optimizer, static_mask, dynamic_mask)[0]


def prune_points(self, static_mask, dynamic_mask) -> None:
    """Drop-in for CGaussianModel.prune_points (scene/c_gaussian_model.py:715-763): the masks select the rows to REMOVE.
    Parameters, optimizer moments and the statistics tensors are compacted by one launch."""
    valid_static = ~static_mask
    if dynamic_mask.shape[0] == 0:
        dynamic_mask = torch.empty(0, dtype=torch.bool, device=static_mask.device)
    valid_dynamic = ~dynamic_mask
    has_dynamic = valid_dynamic.shape[0] != 0
    opt, stat_s, stat_d = _prune(self.optimizer, valid_static, valid_dynamic, [getattr(self, n) for n in STATIC_STATS],
                                 [getattr(self, n) for n in DYNAMIC_STATS] if has_dynamic else [])
    for name, attr in _STATIC_ATTR.items():
        setattr(self, attr, opt[name])
    for n, t in zip(STATIC_STATS, stat_s):
        setattr(self, n, t)
    if not has_dynamic:
        return
    for name, attr in _DYNAMIC_ATTR.items():
        setattr(self, attr, opt[name])
    for n, t in zip(DYNAMIC_STATS, stat_d):
        setattr(self, n, t)


def cat_tensors_to_optimizer(self, tensors_dict) -> Dict[str, nn.Parameter]:
    """Drop-in for CGaussianModel.cat_tensors_to_optimizer (scene/c_gaussian_model.py:765-787): every named group grows by
    its extension tensor, its moments by as many zero rows; groups without an entry are skipped, as in the reference."""
    optimizer = self.optimizer
    jobs = _Jobs()
    planned = []
    for group in optimizer.param_groups:
        assert len(group["params"]) == 1
        if group["name"] not in tensors_dict.keys():
            continue
        ext = tensors_dict[group["name"]]
        p = group["params"][0]
        n, e = int(p.shape[0]), int(ext.shape[0])
        st = optimizer.state.get(p, None)
        new_p = jobs.add(p, n + e, n_a=n, b=ext)
        new_avg = jobs.add(st["exp_avg"], n + e, n_a=n) if st is not None else None
        new_sq = jobs.add(st["exp_avg_sq"], n + e, n_a=n) if st is not None else None
        planned.append((group, new_p, new_avg, new_sq))
    jobs.launch()
    optimizable_tensors = {}
    for group, new_p, new_avg, new_sq in planned:
        optimizable_tensors[group["name"]] = _rebuild_group(optimizer, group, new_p, new_avg, new_sq)
    return optimizable_tensors


def install(model) -> None:
    """Bind the three methods onto a reference CGaussianModel instance (or any object with its attribute names and an
    `optimizer`): everything the class does on top of them keeps working unchanged."""
    model._prune_optimizer = types.MethodType(_prune_optimizer, model)
    model.prune_points = types.MethodType(prune_points, model)
    model.cat_tensors_to_optimizer = types.MethodType(cat_tensors_to_optimizer, model)
