"""Plain-PyTorch restatement of the model-side getters of CGaussianModel - TEST INFRASTRUCTURE.

get_xyz_at_t / get_rotation_at_t / get_scaling / get_opacity_at_t (scene/c_gaussian_model.py:170-215,330-375) and the
interpolation helpers they call (utils/interpolations.py:33-61,81-93), restated in float32 PyTorch on the tensors of an
ex4dgs_b200.synth.Scene.  Pinned bit for bit against the reference's own utils/interpolations.py by
tests/test_synth.py (tests/golden/interp_fixture.npz).  Used to build the pre-interpolated [P, .] tensors that the
unchanged-API path receives in the parity tests, and as the torch reference of the fused front-end kernels
(tests/test_gpu_frontend.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs import this.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ex4dgs_b200.synth import SEED, Scene


def cube_interp(y: torch.Tensor, k: int, d: float) -> torch.Tensor:
    """interpolations.py:81-93 on keyframes k-1..k+2 (c_gaussian_model.py:118)."""
    h00 = 2 * d ** 3 - 3 * d ** 2 + 1
    h10 = d ** 3 - 2 * d ** 2 + d
    h01 = -2 * d ** 3 + 3 * d ** 2
    h11 = d ** 3 - d ** 2
    ykm1, yk, yk1, yk2 = y[:, k - 1], y[:, k], y[:, k + 1], y[:, k + 2]
    return h00 * yk + h10 * ((yk1 - ykm1) / 2) + h01 * yk1 + h11 * ((yk2 - yk) / 2)


def quat_slerp(v1: torch.Tensor, v2: torch.Tensor, t: float) -> torch.Tensor:
    """interpolations.py:33-52 (no shortest-path flip)."""
    v1 = v1 / torch.norm(v1, dim=-1, keepdim=True)
    v2 = v2 / torch.norm(v2, dim=-1, keepdim=True)
    d = (v1 * v2).sum(-1, keepdim=True).clamp(-1 + 1e-4, 1 - 1e-4)
    omega = torch.acos(d).clamp_min(1e-4)
    s_omega = torch.sin(omega).clamp_min(1e-4)
    p0 = torch.sin((1 - t) * omega) / s_omega
    p1 = torch.sin(t * omega) / s_omega
    ps = (p0 + p1).clamp_min(1e-4)
    p0, p1 = p0 / ps, p1 / ps
    ret = v1 * p0 + v2 * p1
    ret = torch.where(ret.abs().sum(-1, keepdim=True) > 1e-4, ret, v1)
    return ret / ret.norm(dim=-1, keepdim=True)


def time_bigaussian(mean: torch.Tensor, var: torch.Tensor, t: float, var_min: float) -> torch.Tensor:
    """interpolations.py:55-61."""
    m = (t - mean).min(dim=1)[0]
    v = torch.where((t > mean).any(dim=1), var[:, 1], var[:, 0])
    o = torch.exp(-1 * (m.pow(2) / (v.exp() + var_min / 2.36).pow(2)))
    return torch.where((mean[:, 0] - t) * (mean[:, 1] - t) < 0, torch.ones_like(o), o)


def frame_indices(sc: Scene, t: Optional[float] = None):
    t = sc.timestamp if t is None else t
    tt = t + sc.time_shift
    k = int(tt // sc.interval)
    d = (tt % sc.interval) / sc.interval
    return k, d


def model_getters(sc, t: Optional[float] = None):
    """get_xyz_at_t / get_rotation_at_t / get_scaling / get_opacity_at_t of CGaussianModel
    (scene/c_gaussian_model.py:170-215,330-375) in plain differentiable PyTorch: static first, then dynamic."""
    t = sc.timestamp if t is None else t
    k, d = frame_indices(sc, t)
    nd = sc.xyz_motion.shape[0]
    means_s = sc.xyz + sc.xyz_disp * t / sc.duration
    if nd:
        means_d = cube_interp(sc.xyz_motion, k, d)
        rot_d = quat_slerp(sc.rotation_motion[:, k], sc.rotation_motion[:, k + 1], d)
        tau = (t + sc.time_shift) / sc.interval
        op_d = (time_bigaussian(sc.opacity_center, sc.opacity_var, tau, sc.var_pad / sc.interval)[:, None]
                * torch.sigmoid(sc.opacity_motion))
        means = torch.cat([means_s, means_d]).contiguous()
        rots = torch.cat([sc.rotation, rot_d]).contiguous()
        opac = torch.cat([torch.sigmoid(sc.opacity), op_d]).contiguous()
        scales = torch.exp(torch.cat([sc.scaling, sc.scaling_motion])).contiguous()
    else:
        means, rots, opac = means_s.contiguous(), sc.rotation, torch.sigmoid(sc.opacity)
        scales = torch.exp(sc.scaling)
    return means, rots, scales, opac


def flat_inputs(sc: Scene, t: Optional[float] = None) -> Dict[str, torch.Tensor]:
    """What gaussian_renderer/__init__.py:62-95 hands to the rasterizer: static first, then dynamic."""
    means, rots, scales, opac = model_getters(sc, t)
    if sc.xyz_motion.shape[0]:
        shs = torch.cat([sc.features, sc.features_motion]).contiguous()
    else:
        shs = sc.features
    P = means.shape[0]
    if getattr(sc, "_dir_nonzero", False):
        g = torch.Generator().manual_seed(getattr(sc, "_seed", SEED) + 1)
        dir3d = 0.3 * torch.randn(P, 3, generator=g)
    else:
        dir3d = torch.zeros(P, 3)
    return dict(means3D=means.float(), dir3D=dir3d, opacities=opac.float(), shs=shs.float(),
                scales=scales.float(), rotations=rots.float())


