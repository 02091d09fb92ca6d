"""Fused front-end (row N1) against the plain-PyTorch fp32 restatement of the reference's getters
(ex4dgs_b200/synth.py, itself checked bit-for-bit against utils/interpolations.py in test_synth.py):
forward values and every gradient via torch.autograd."""
import numpy as np
import pytest
import torch

from ex4dgs_b200 import synth

pytestmark = pytest.mark.gpu


def _torch_getters(sc, T, t):
    """differentiable PyTorch path, same math as c_gaussian_model.py:170-215,330-375"""
    k, d = synth.frame_indices(sc, t)
    means_s = T["xyz"] + T["xyz_disp"] * t / sc.duration
    means_d = synth.cube_interp(T["xyz_motion"], k, d)
    rot_d = synth.quat_slerp(T["rotation_motion"][:, k], T["rotation_motion"][:, k + 1], d)
    tau = (t + sc.time_shift) / sc.interval
    op_d = synth.time_bigaussian(T["opacity_center"], T["opacity_var"], tau, sc.var_pad / sc.interval)[:, None] * \
        torch.sigmoid(T["opacity_motion"])
    return (torch.cat([means_s, means_d]), torch.cat([T["rotation"], rot_d]),
            torch.exp(torch.cat([T["scaling"], T["scaling_motion"]])), torch.cat([torch.sigmoid(T["opacity"]), op_d]))


@pytest.mark.parametrize("t", [0.0, 41.5, 137.0, 299.0])
def test_frontend_matches_torch(built, t):
    from ex4dgs_b200.frontend import interpolate_gaussians
    sc = synth.make_scene(700, 900, 64, 48, seed=31)
    names = ["xyz", "xyz_disp", "rotation", "scaling", "opacity", "xyz_motion", "rotation_motion", "scaling_motion",
             "opacity_motion", "opacity_center", "opacity_var"]
    A = {n: getattr(sc, n).clone().cuda().requires_grad_(True) for n in names}
    B = {n: getattr(sc, n).clone().cuda().requires_grad_(True) for n in names}
    out_a = interpolate_gaussians(*[A[n] for n in names], t=t, duration=sc.duration, interval=sc.interval,
                                  time_shift=sc.time_shift, var_min=sc.var_pad / sc.interval)
    out_b = _torch_getters(sc, B, t)
    g = torch.Generator().manual_seed(3)
    ups = [torch.randn(o.shape, generator=g).cuda() for o in out_b]
    for oa, ob, nm in zip(out_a, out_b, ["means", "rot", "scales", "opac"]):
        assert oa.shape == ob.shape
        assert torch.allclose(oa, ob, rtol=2e-5, atol=2e-6), (nm, float((oa - ob).abs().max()))
    torch.autograd.backward(list(out_a), ups)
    torch.autograd.backward(list(out_b), ups)
    for n in names:
        ga, gb = A[n].grad, B[n].grad
        assert ga is not None and gb is not None, n
        scale = max(1e-6, float(gb.abs().max()))
        err = float((ga - gb).abs().max()) / scale
        assert err <= 2e-4, (n, err)


def test_frontend_feeds_rasterizer(built):
    """fused front-end + rasterizer == PyTorch getters + rasterizer (image within 1e-4)"""
    import ex4dgs_b200 as m
    from ex4dgs_b200.frontend import interpolate_gaussians
    from tests import _util as U
    sc = synth.make_config("C1d")
    names = ["xyz", "xyz_disp", "rotation", "scaling", "opacity", "xyz_motion", "rotation_motion", "scaling_motion",
             "opacity_motion", "opacity_center", "opacity_var"]
    T = {n: getattr(sc, n).cuda() for n in names}
    means, rots, scales, opac = interpolate_gaussians(*[T[n] for n in names], t=sc.timestamp, duration=sc.duration,
                                                      interval=sc.interval, time_shift=sc.time_shift,
                                                      var_min=sc.var_pad / sc.interval)
    shs = torch.cat([sc.features, sc.features_motion]).cuda()
    rs = U.settings_for(m, sc, "cuda")
    z = torch.zeros_like(means)
    img = m.GaussianRasterizer(rs)(means3D=means, means2D=z, dir3D=z, opacities=opac, shs=shs, scales=scales, rotations=rots)[0]
    ref = U.run_impl(m, sc, kind="ours", grads=False, intermediates=False)["color"]
    assert float(np.abs(img.cpu().numpy() - ref).max()) <= 1e-4
