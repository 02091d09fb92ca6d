"""Fused front-end (row N1) against the plain-PyTorch fp32 restatement of the reference's getters
(ex4dgs_b200/synth.py, itself checked bit-for-bit against utils/interpolations.py in test_synth.py):
forward values and every gradient via torch.autograd."""
import numpy as np
import pytest
import torch

from ex4dgs_b200 import synth
from oracle import getters_oracle as GO  # noqa: E402

pytestmark = pytest.mark.gpu


def _torch_getters(sc, T, t):
    """differentiable PyTorch path, same math as c_gaussian_model.py:170-215,330-375"""
    k, d = GO.frame_indices(sc, t)
    means_s = T["xyz"] + T["xyz_disp"] * t / sc.duration
    means_d = GO.cube_interp(T["xyz_motion"], k, d)
    rot_d = GO.quat_slerp(T["rotation_motion"][:, k], T["rotation_motion"][:, k + 1], d)
    tau = (t + sc.time_shift) / sc.interval
    op_d = GO.time_bigaussian(T["opacity_center"], T["opacity_var"], tau, sc.var_pad / sc.interval)[:, None] * \
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


def _live_model(sc, cls):
    """The reference's UNMODIFIED CGaussianModel (oracle/_ref/callers/scene/c_gaussian_model.py), built by its own
    constructor, holding the synthetic scene in nn.Parameters of the reference's shapes (_opacity_duration_* are [Nd,2,1])."""
    import torch.nn as nn
    m = cls(sh_degree=3, duration=int(sc.duration), interval=int(sc.interval), time_pad=int(sc.time_pad), interp_type="cube",
            rot_interp_type="slerp", var_pad=sc.var_pad, kernel_size=sc.cam.kernel_size)
    P = lambda t: nn.Parameter(t.detach().clone().cuda().requires_grad_(True))     # noqa: E731
    m._xyz, m._xyz_disp, m._rotation, m._scaling, m._opacity = P(sc.xyz), P(sc.xyz_disp), P(sc.rotation), P(sc.scaling), P(sc.opacity)
    m._features_dc, m._features_rest = P(sc.features[:, :1]), P(sc.features[:, 1:])
    m._xyz_motion, m._rotation_motion = P(sc.xyz_motion), P(sc.rotation_motion)
    m._scaling_motion, m._opacity_motion = P(sc.scaling_motion), P(sc.opacity_motion)
    m._opacity_duration_center, m._opacity_duration_var = P(sc.opacity_center[:, :, None]), P(sc.opacity_var[:, :, None])
    m._features_dc_motion, m._features_rest_motion = P(sc.features_motion[:, :1]), P(sc.features_motion[:, 1:])
    m.active_sh_degree = 3
    return m


_PARAMS = ("_xyz", "_xyz_disp", "_rotation", "_scaling", "_opacity", "_xyz_motion", "_rotation_motion", "_scaling_motion",
           "_opacity_motion", "_opacity_duration_center", "_opacity_duration_var", "_features_dc", "_features_rest",
           "_features_dc_motion", "_features_rest_motion")


# both ends of the range the reference asserts (c_gaussian_model.py:171: -time_shift <= t <= duration + time_shift would
# read keyframe -1; the trainer's timestamps are 0 .. duration), keyframe boundaries (delta = 0) and interior points
@pytest.mark.parametrize("t", [0.0, 8.0, 41.5, 137.0, 290.0, 300.0])
def test_fused_getters_against_live_reference_class(built, t):
    """FusedGetters(model) against the model's OWN getters - get_xyz_at_t / get_rotation_at_t / get_opacity_at_t /
    get_scaling / get_features of the reference's CGaussianModel (scene/c_gaussian_model.py:170-215,330-375), on a real
    instance of that class: values and the gradients that reach every nn.Parameter."""
    import bench
    from ex4dgs_b200.frontend import FusedGetters
    cls = bench.load_reference_model_class()
    if cls is None:
        pytest.skip("oracle/_ref/callers not installed (no /root/reference at build time)")
    sc = synth.make_scene(700, 900, 64, 48, seed=31)
    ref, ours = _live_model(sc, cls), _live_model(sc, cls)
    assert tuple(ref._opacity_duration_center.shape) == (900, 2, 1) and ref.time_shift == sc.time_shift
    fg = FusedGetters(ours)
    a = (fg.get_xyz_at_t(t), fg.get_rotation_at_t(t), fg.get_scaling(), fg.get_opacity_at_t(t), fg.get_features().cat())
    b = (ref.get_xyz_at_t(t), ref.get_rotation_at_t(t), ref.get_scaling(), ref.get_opacity_at_t(t), ref.get_features())
    g = torch.Generator().manual_seed(5)
    ups = [torch.randn(o.shape, generator=g).cuda() for o in b]
    for oa, ob, nm in zip(a, b, ["xyz", "rotation", "scaling", "opacity", "features"]):
        assert oa.shape == ob.shape, nm
        assert torch.allclose(oa, ob, rtol=2e-5, atol=2e-6), (nm, float((oa - ob).abs().max()))
    torch.autograd.backward(list(a), ups)
    torch.autograd.backward(list(b), ups)
    for n in _PARAMS:
        ga, gb = getattr(ours, n).grad, getattr(ref, n).grad
        assert ga is not None and gb is not None and ga.shape == gb.shape, n
        err = float((ga - gb).abs().max()) / max(1e-6, float(gb.abs().max()))
        assert err <= 2e-4, (n, err)


def test_fused_getters_cache_follows_the_parameters(built):
    """A FusedGetters wrapper kept across iterations: optimizer steps (FusedRAdam writes through raw pointers and bumps
    the version counters), swapped-in Parameters (densification, reset_opacity) and a second backward at the same
    timestamp must all see fresh values / a fresh graph."""
    import types
    from ex4dgs_b200.frontend import FusedGetters
    from ex4dgs_b200.optim import FusedRAdam
    import torch.nn as nn
    sc = synth.make_scene(300, 200, 64, 48, seed=7)
    P = lambda x: nn.Parameter(x.detach().clone().cuda())      # noqa: E731
    m = types.SimpleNamespace(
        _xyz=P(sc.xyz), _xyz_disp=P(sc.xyz_disp), _rotation=P(sc.rotation), _scaling=P(sc.scaling), _opacity=P(sc.opacity),
        _xyz_motion=P(sc.xyz_motion), _rotation_motion=P(sc.rotation_motion), _scaling_motion=P(sc.scaling_motion),
        _opacity_motion=P(sc.opacity_motion), _opacity_duration_center=P(sc.opacity_center[:, :, None]),
        _opacity_duration_var=P(sc.opacity_var[:, :, None]), _features_dc=P(sc.features[:, :1]),
        _features_rest=P(sc.features[:, 1:]), _features_dc_motion=P(sc.features_motion[:, :1]),
        _features_rest_motion=P(sc.features_motion[:, 1:]),
        duration=sc.duration, interval=sc.interval, time_shift=sc.time_shift, var_pad=sc.var_pad)
    fg = FusedGetters(m)
    t = 137.0
    x0 = fg.get_xyz_at_t(t)
    assert fg.get_xyz_at_t(t) is x0                       # one launch serves the getters of a frame
    s0 = fg.get_scaling().detach().clone()
    x0.sum().backward()
    # same timestamp, graph consumed: the next frame (after get_features) must build a new one
    fg.get_features()
    x1 = fg.get_xyz_at_t(t)
    assert x1 is not x0
    (x1.sum() + fg.get_scaling().sum()).backward()
    # optimizer step through raw pointers -> new values
    opt = FusedRAdam([{"params": [m._scaling], "lr": 0.5, "name": "scaling"}, {"params": [m._xyz], "lr": 0.5, "name": "xyz"}])
    v0 = m._scaling._version
    opt.step()
    assert m._scaling._version > v0
    x2 = fg.get_xyz_at_t(t)
    assert x2 is not x1 and not torch.equal(fg.get_scaling().detach(), s0)
    # a swapped-in Parameter of another size (densification) with a version counter starting again at 0
    m._xyz = P(sc.xyz[:100])
    m._xyz_disp, m._rotation, m._scaling, m._opacity = P(sc.xyz_disp[:100]), P(sc.rotation[:100]), P(sc.scaling[:100]), P(sc.opacity[:100])
    assert fg.get_xyz_at_t(t).shape[0] == 100 + 200
    # no_grad evaluation at the same timestamp does not reuse the recording result either
    with torch.no_grad():
        assert fg.get_xyz_at_t(t).requires_grad is False
