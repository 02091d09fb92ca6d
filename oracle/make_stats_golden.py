"""Generates tests/golden/stats_fixture.npz by running the reference's OWN bookkeeping code on seeded inputs:
CGaussianModel.mark_prune_stats / add_densification_stats / add_l1_ssim_stats
(/root/reference/scene/c_gaussian_model.py:1095-1145, the class imported unmodified and its methods called
on an instance that only carries the statistics tensors), the two max_radii2D lines of train.py:205-206
(quoted verbatim below) and the regularisation expressions of train.py:156-162 under autograd.  Everything
is pure PyTorch and runs on the CPU.  Test infrastructure; build container only:

    python oracle/make_stats_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_ply_golden import _install_stubs  # noqa: E402
from oracle.stats_oracle import ALL_NAMES  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "stats_fixture.npz")


def make_inputs(Ns, Nd, seed, steps):
    """Seeded per-iteration inputs shaped like what the training loop sees: ~45 % invisible Gaussians
    (radius 0), error gradients that are exactly 0 for a third of the visible ones, a few tiny / negative
    accumulated-alpha values around the 1e-4 clamp and the 0.01 threshold."""
    g = torch.Generator().manual_seed(seed)
    P = Ns + Nd
    its = []
    for s in range(steps):
        radii = torch.randint(1, 60, (P,), generator=g, dtype=torch.int32)
        radii[torch.rand(P, generator=g) < 0.45] = 0
        grad = torch.randn(P, 3, generator=g) * 1e-3
        err = torch.rand(P, 3, generator=g)
        err[:, 0] = err[:, 0] * torch.tensor([1.0, 0.02, 2e-4])[torch.randint(0, 3, (P,), generator=g)]
        err[torch.rand(P, generator=g) < 0.3] = 0.0
        err[radii == 0] = 0.0                 # invisible Gaussians receive no back-projected error
        odd = torch.rand(P, generator=g) < 0.01
        err[odd, 0] = -err[odd, 0]            # acc back-projection is >= 0 in practice; the code must not care
        its.append(dict(radii=radii, grad=grad, err=err, timestamp=float(7 + 13 * s), densify=(s != steps - 1)))
    return its


def fresh_state(Ns, Nd):
    """CGaussianModel.training_setup (c_gaussian_model.py:412-428) and :408-409 / :843-844, on the CPU."""
    z, o = torch.zeros, torch.ones
    return dict(
        max_radii2D=z(Ns), min_radii2D=o(Ns) * 1000, xyz_gradient_accum=z(Ns, 1), denom=z(Ns, 1), xyz_error_accum=z(Ns, 1),
        xyz_error_min=o(Ns, 1) * 1000, xyz_error_min_timestamp=o(Ns, 1) * -1, xyz_ssim_error_accum=z(Ns, 1), error_denom=z(Ns, 1),
        motion_max_radii2D=z(Nd), motion_min_radii2D=o(Nd) * 1000, motion_xyz_gradient_accum=z(Nd, 1), motion_denom=z(Nd, 1),
        motion_xyz_error_min=o(Nd, 1) * 1000, motion_xyz_error_mean=z(Nd, 1), motion_xyz_error_min_timestamp=o(Nd, 1) * -1,
        motion_xyz_ssim_error_accum=z(Nd, 1), motion_error_denom=z(Nd, 1))


def main():
    _install_stubs()
    sys.path.insert(0, "/root/reference")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_c_gaussian_model", "/root/reference/scene/c_gaussian_model.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    CGaussianModel = mod.CGaussianModel

    out = {}
    for tag, (Ns, Nd, l1_accum) in {"a": (301, 77, True), "b": (64, 0, True), "c": (50, 33, False)}.items():
        gaussians = CGaussianModel.__new__(CGaussianModel)
        for k, v in fresh_state(Ns, Nd).items():
            setattr(gaussians, k, v)
        gaussians._xyz = torch.zeros(Ns, 3)
        its = make_inputs(Ns, Nd, 100 + ord(tag), steps=4)
        for it in its:
            radii = it["radii"]
            vp = types.SimpleNamespace(grad=it["grad"])              # viewspace_point_tensor
            ve = types.SimpleNamespace(grad=it["err"])               # viewspace_point_error_tensor
            visibility_filter = radii > 0                            # gaussian_renderer/__init__.py:123
            # ---- train.py:196-212, verbatim ----
            if l1_accum:
                gaussians.mark_prune_stats(radii, ve)
            if it["densify"]:
                static_num = gaussians._xyz.shape[0]
                static_vis_filter = visibility_filter[:static_num]
                static_radii = radii[:static_num]
                dynamic_vis_filter = visibility_filter[static_num:]
                dynamic_radii = radii[static_num:]
                gaussians.max_radii2D[static_vis_filter] = torch.max(gaussians.max_radii2D[static_vis_filter], static_radii[static_vis_filter])
                gaussians.motion_max_radii2D[dynamic_vis_filter] = torch.max(gaussians.motion_max_radii2D[dynamic_vis_filter], dynamic_radii[dynamic_vis_filter])
                gaussians.add_densification_stats(vp, static_vis_filter, dynamic_vis_filter, static_num)
                if l1_accum:
                    gaussians.add_l1_ssim_stats(ve, static_vis_filter, dynamic_vis_filter, static_num, it["timestamp"])
        out["%s_meta" % tag] = np.array([Ns, Nd, int(l1_accum), 100 + ord(tag), 4], dtype=np.int64)
        for k in ALL_NAMES:
            out["%s_%s" % (tag, k)] = getattr(gaussians, k).numpy()

    # regularisers: train.py:156-162, float32 exactly as the training loop evaluates them
    g = torch.Generator().manual_seed(4242)
    Ns, Nd, K = 257, 41, 9
    disp = (torch.randn(Ns, 3, generator=g) * 0.05)
    disp[5] = 0.0                                                    # norm backward at 0 -> 0
    motion = torch.randn(Nd, K, 3, generator=g)
    motion[3, 4] = motion[3, 0]                                      # a zero difference
    static_reg, motion_reg = 0.0001, 0.0001                          # arguments/__init__.py:134-135
    _xyz_disp = disp.clone().requires_grad_(True)
    _xyz_motion = motion.clone().requires_grad_(True)
    loss = torch.zeros(())
    t0 = static_reg * torch.log(_xyz_disp.norm(dim=-1) + 0.001).mean()
    diff1 = (_xyz_motion[:, :1] - _xyz_motion[:, 1:])
    t1 = motion_reg * diff1.norm(dim=-1).mean()
    loss = loss + t0 + t1
    loss.backward()
    out.update(reg_disp=disp.numpy(), reg_motion=motion.numpy(), reg_terms=np.array([t0.item(), t1.item()], dtype=np.float32),
               reg_gdisp=_xyz_disp.grad.numpy(), reg_gmotion=_xyz_motion.grad.numpy(),
               reg_weights=np.array([static_reg, motion_reg], dtype=np.float64))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
