"""CPU restatement of the reference's per-iteration bookkeeping and regularisation terms (TEST
INFRASTRUCTURE: only tests/ and bench.py's baseline legs may import this; the product never does).

iteration_stats() follows, update by update and in the reference's order,
    train.py:196-197            if opt.l1_accum: gaussians.mark_prune_stats(radii, viewspace_point_error_tensor)
    train.py:200-212            max_radii2D / motion_max_radii2D, add_densification_stats, add_l1_ssim_stats
    scene/c_gaussian_model.py:1095-1103 (add_densification_stats), :1105-1117 (mark_prune_stats),
    :1119-1145 (add_l1_ssim_stats)
with dense torch.where arithmetic instead of boolean-mask indexing (same values: every masked update of
the reference is element-wise).  regularizers() is the literal text of train.py:156-162 under autograd.
Pinned by tests/golden/stats_fixture.npz, which oracle/make_stats_golden.py produces by calling the
reference's UNMODIFIED CGaussianModel methods and the train.py expressions.
"""
import torch

STATIC = dict(max_radii2D="max_radii2D", min_radii2D="min_radii2D", grad_accum="xyz_gradient_accum", denom="denom",
              err_accum="xyz_error_accum", err_min="xyz_error_min", err_min_ts="xyz_error_min_timestamp",
              ssim_accum="xyz_ssim_error_accum", err_denom="error_denom")
DYNAMIC = dict(max_radii2D="motion_max_radii2D", min_radii2D="motion_min_radii2D", grad_accum="motion_xyz_gradient_accum",
               denom="motion_denom", err_accum="motion_xyz_error_mean", err_min="motion_xyz_error_min",
               err_min_ts="motion_xyz_error_min_timestamp", ssim_accum="motion_xyz_ssim_error_accum",
               err_denom="motion_error_denom")
ALL_NAMES = list(STATIC.values()) + list(DYNAMIC.values())


def _part(state, names, radii, grad, err, timestamp, densify):
    rf = radii.to(torch.float32)
    vis = radii > 0
    if err is not None:                                           # mark_prune_stats
        f = err[:, 0] > 0
        m = state[names["min_radii2D"]]
        state[names["min_radii2D"]] = torch.where(f, torch.min(m, rf), m)
    if not densify:
        return
    m = state[names["max_radii2D"]]
    state[names["max_radii2D"]] = torch.where(vis, torch.max(m, rf), m)
    v1 = vis.unsqueeze(1)
    state[names["grad_accum"]] = state[names["grad_accum"]] + torch.where(v1, torch.norm(grad[:, :2], dim=-1, keepdim=True), torch.zeros(1))
    state[names["denom"]] = state[names["denom"]] + v1.float()
    if err is None:
        return
    e0 = err[:, 0:1]
    l1 = err[:, 1:2] / e0.clamp_min(1e-4)
    better = torch.logical_and(state[names["err_min"]] > l1, e0 > 0.01) & v1
    state[names["err_accum"]] = state[names["err_accum"]] + torch.where(v1, l1, torch.zeros(1))
    state[names["err_min_ts"]] = torch.where(better, torch.full_like(l1, float(timestamp)), state[names["err_min_ts"]])
    state[names["err_min"]] = torch.where(better, l1, state[names["err_min"]])
    state[names["ssim_accum"]] = state[names["ssim_accum"]] + torch.where(v1, err[:, 2:3] / e0.clamp_min(1e-4), torch.zeros(1))
    state[names["err_denom"]] = state[names["err_denom"]] + ((e0 > 0) & v1).float()


def iteration_stats(state, Ns, radii, grad_means2D, grad_error, timestamp, densify=True):
    """state: dict attribute name -> CPU float tensor (the model's shapes: [N] radii arrays, [N,1] others);
    updated in place (entries are replaced)."""
    _part(state, STATIC, radii[:Ns], grad_means2D[:Ns], None if grad_error is None else grad_error[:Ns], timestamp, densify)
    if radii.shape[0] > Ns:
        _part(state, DYNAMIC, radii[Ns:], grad_means2D[Ns:], None if grad_error is None else grad_error[Ns:], timestamp, densify)
    return state


def regularizers(xyz_disp, xyz_motion, static_reg, motion_reg, dtype=torch.float64):
    """(terms [2], dL/dxyz_disp, dL/dxyz_motion) of train.py:156-162, evaluated in `dtype`."""
    d = xyz_disp.detach().to(dtype).requires_grad_(True)
    m = xyz_motion.detach().to(dtype).requires_grad_(True)
    terms = [torch.zeros((), dtype=dtype), torch.zeros((), dtype=dtype)]
    if static_reg > 0 and d.shape[0] > 0:
        terms[0] = static_reg * torch.log(d.norm(dim=-1) + 0.001).mean()
    if motion_reg > 0 and m.shape[0] > 0 and m.shape[1] > 1:
        diff1 = (m[:, :1] - m[:, 1:])
        terms[1] = motion_reg * diff1.norm(dim=-1).mean()
    total = terms[0] + terms[1]
    if total.requires_grad:
        total.backward()
    gd = d.grad if d.grad is not None else torch.zeros_like(d)
    gm = m.grad if m.grad is not None else torch.zeros_like(m)
    return torch.stack([t.detach() for t in terms]), gd, gm
