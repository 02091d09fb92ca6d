"""Fused model front-end: CGaussianModel's per-frame getters in one kernel.

`interpolate_gaussians(...)` returns the flat `[P,.]` tensors that gaussian_renderer/__init__.py:62-81
obtains from `pc.get_xyz_at_t(t)`, `pc.get_rotation_at_t(t)`, `pc.get_scaling()`, `pc.get_opacity_at_t(t)`
(scene/c_gaussian_model.py:170-215,330-375), static Gaussians first, differentiable w.r.t. the
model's native tensors.  interp_type "cube" and rot_interp_type "slerp" (the defaults of every
reference config) are implemented; anything else must keep using the PyTorch getters.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().float().contiguous()


def _p(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


class _FusedFrontEnd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, xyz_disp, rotation, scaling, opacity, xyz_motion, rotation_motion, scaling_motion,
                opacity_motion, opacity_center, opacity_var, t, duration, interval, time_shift, var_min):
        if not xyz.is_cuda:
            raise RuntimeError("ex4dgs_b200: the fused front-end is CUDA-only")
        lib = _lib.load()
        dev = xyz.device
        Ns, Nd = xyz.shape[0], xyz_motion.shape[0]
        K = xyz_motion.shape[1] if Nd else 0
        ten = [_c(x) for x in (xyz, xyz_disp, rotation, scaling, opacity, xyz_motion, rotation_motion, scaling_motion,
                               opacity_motion, opacity_center, opacity_var)]
        P = Ns + Nd
        means = torch.empty(P, 3, device=dev)
        rots = torch.empty(P, 4, device=dev)
        scales = torch.empty(P, 3, device=dev)
        opac = torch.empty(P, 1, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.ex4dgs_frontend_forward(Ns, Nd, K, *[_p(x) for x in ten], float(t), float(duration), float(interval),
                                             float(time_shift), float(var_min), _p(means), _p(rots), _p(scales), _p(opac),
                                             C.c_void_p(stream))
        if rc < 0:
            raise RuntimeError("ex4dgs_frontend_forward failed (%d): timestamp outside the keyframe range?" % rc)
        ctx.save_for_backward(*ten)
        ctx.scalars = (float(t), float(duration), float(interval), float(time_shift), float(var_min))
        ctx.shapes = [x.shape for x in (xyz, xyz_disp, rotation, scaling, opacity, xyz_motion, rotation_motion,
                                        scaling_motion, opacity_motion, opacity_center, opacity_var)]
        return means, rots, scales, opac

    @staticmethod
    def backward(ctx, g_means, g_rots, g_scales, g_opac):
        lib = _lib.load()
        (xyz, xyz_disp, rotation, scaling, opacity, xyz_motion, rotation_motion, scaling_motion, opacity_motion,
         opacity_center, opacity_var) = ctx.saved_tensors
        dev = xyz.device
        Ns, Nd = xyz.shape[0], xyz_motion.shape[0]
        K = xyz_motion.shape[1] if Nd else 0
        outs = [torch.empty_like(x) for x in (xyz, xyz_disp, rotation, scaling, opacity, xyz_motion, rotation_motion,
                                              scaling_motion, opacity_motion, opacity_center, opacity_var)]
        g = [_c(x) for x in (g_means, g_rots, g_scales, g_opac)]
        stream = torch.cuda.current_stream(dev).cuda_stream
        t, duration, interval, time_shift, var_min = ctx.scalars
        with torch.cuda.device(dev):
            rc = lib.ex4dgs_frontend_backward(Ns, Nd, K, _p(rotation_motion), _p(scaling), _p(opacity), _p(scaling_motion),
                                              _p(opacity_motion), _p(opacity_center), _p(opacity_var),
                                              t, duration, interval, time_shift, var_min,
                                              *[_p(x) for x in g], *[_p(x) for x in outs], C.c_void_p(stream))
        if rc < 0:
            raise RuntimeError("ex4dgs_frontend_backward failed (%d)" % rc)
        outs = [o.reshape(s) for o, s in zip(outs, ctx.shapes)]
        return (*outs, None, None, None, None, None)


def interpolate_gaussians(xyz, xyz_disp, rotation, scaling, opacity, xyz_motion, rotation_motion, scaling_motion,
                          opacity_motion, opacity_center, opacity_var, *, t, duration, interval, time_shift, var_min):
    """-> (means3D [P,3], rotations [P,4], scales [P,3], opacities [P,1]); static Gaussians first."""
    return _FusedFrontEnd.apply(xyz, xyz_disp, rotation, scaling, opacity, xyz_motion, rotation_motion, scaling_motion,
                                opacity_motion, opacity_center, opacity_var, t, duration, interval, time_shift, var_min)


_GETTER_INPUTS = ("_xyz", "_xyz_disp", "_rotation", "_scaling", "_opacity", "_xyz_motion", "_rotation_motion",
                  "_scaling_motion", "_opacity_motion", "_opacity_duration_center", "_opacity_duration_var")


class FusedGetters:
    """Duck-typed stand-in for the per-frame getters of CGaussianModel: wrap the model and hand the
    wrapper to the reference's unmodified render() (gaussian_renderer/__init__.py:28,62-95).  The four
    getters that render() calls for one timestamp (get_xyz_at_t twice, get_opacity_at_t,
    get_scaling, get_rotation_at_t) are served from ONE fused kernel launch; get_features returns the model's
    four SH tensors as a SegmentedSH (no torch.cat: render() only forwards that object to the rasterizer, which
    reads the tensors in place and writes their gradients directly).  Every other attribute is forwarded to the model.

    Cache: the result of the fused launch is kept for the getters of ONE frame.  It is dropped by get_features()
    (the last getter render() calls) and is only reused while its key is unchanged - the timestamp, grad mode, and
    identity, storage, shape and version counter of all eleven input tensors (so densification / reset_opacity, which
    swap in new Parameters, and optimizer steps, which bump the versions - FusedRAdam does so explicitly -
    invalidate it).  A wrapper may therefore be kept across iterations; a caller that drives the getters itself and
    never calls get_features() should call new_frame() between frames when autograd is recording (a cached result
    carries the graph of the frame it was computed in).

    `model` needs the reference's attribute names: _xyz, _xyz_disp, _rotation, _scaling, _opacity,
    _xyz_motion, _rotation_motion, _scaling_motion, _opacity_motion, _opacity_duration_center,
    _opacity_duration_var, duration, interval, time_shift, var_pad (interp_type "cube", slerp)."""

    def __init__(self, model):
        object.__setattr__(self, "_m", model)
        object.__setattr__(self, "_key", None)
        object.__setattr__(self, "_val", None)

    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, "_m"), name)

    def new_frame(self):
        """Drop the cached result of the fused launch."""
        object.__setattr__(self, "_key", None)
        object.__setattr__(self, "_val", None)

    def _frame(self, t):
        m = self._m
        ten = [getattr(m, n) for n in _GETTER_INPUTS]
        key = (float(t), torch.is_grad_enabled(), float(m.duration), float(m.interval), float(m.time_shift), float(m.var_pad),
               tuple((id(x), x.data_ptr(), tuple(x.shape), int(x._version), bool(x.requires_grad)) for x in ten))
        if self._key != key:
            val = interpolate_gaussians(*ten, t=float(t), duration=float(m.duration), interval=float(m.interval),
                                        time_shift=float(m.time_shift), var_min=float(m.var_pad) / float(m.interval))
            object.__setattr__(self, "_key", key)
            object.__setattr__(self, "_val", val)
        return self._val

    def get_xyz_at_t(self, t, mode=0, training=True):
        assert mode == 0, "fused getters implement mode 0 (static + dynamic)"
        return self._frame(t)[0]

    def get_rotation_at_t(self, t, mode=0):
        assert mode == 0
        return self._frame(t)[1]

    def get_scaling(self, mode=0):
        assert mode == 0
        if self._val is None:
            raise RuntimeError("get_scaling() before any get_*_at_t(t): the fused getters need the timestamp first")
        return self._val[2]

    def get_opacity_at_t(self, t, mode=0, training=False):
        assert mode == 0
        return self._frame(t)[3]

    def get_features(self, mode=0):
        assert mode == 0
        from .rasterizer import SegmentedSH
        m = self._m
        self.new_frame()          # last getter of render(): the next frame interpolates afresh
        return SegmentedSH(m._features_dc, m._features_rest, m._features_dc_motion, m._features_rest_motion)
