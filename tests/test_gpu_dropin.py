"""Drop-in test: the reference's ONLY caller of the rasterizer, gaussian_renderer/__init__.py::render()
(installed unmodified under oracle/_ref/callers by oracle/build_ref.py), runs against this
repository's `diff_gaussian_rasterization_df` package with a stub model / camera, and produces the
same dictionary as the same render() bound to the compiled reference extension."""
import importlib
import math
import os
import sys
import types

import numpy as np
import pytest
import torch

from tests import _util as U
from ex4dgs_b200 import synth
from oracle import getters_oracle as GO  # noqa: E402

pytestmark = pytest.mark.gpu
CALLERS = os.path.join(U.REF_DIR, "callers")


class StubModel:
    """Just the attributes render() touches (gaussian_renderer/__init__.py:28,47,62-95)."""

    def __init__(self, sc):
        self.inp = {k: v.cuda().requires_grad_(True) for k, v in GO.flat_inputs(sc).items()}
        self._xyz = self.inp["means3D"]
        self.kernel_size = sc.cam.kernel_size
        self.active_sh_degree = sc.sh_degree
        self.max_sh_degree = 3

    def get_xyz_at_t(self, t, mode=0, training=True):
        return self.inp["means3D"]

    def get_opacity_at_t(self, t, mode=0, training=False):
        return self.inp["opacities"]

    def get_scaling(self, mode=0):
        return self.inp["scales"]

    def get_rotation_at_t(self, t, mode=0):
        return self.inp["rotations"]

    def get_features(self, mode=0):
        return self.inp["shs"]


def _camera(sc):
    cam = sc.cam
    return types.SimpleNamespace(FoVx=2 * math.atan(cam.tanfovx), FoVy=2 * math.atan(cam.tanfovy), image_height=cam.H,
                                 image_width=cam.W, world_view_transform=cam.viewmatrix.cuda(),
                                 full_proj_transform=cam.projmatrix.cuda(), camera_center=cam.campos.cuda(),
                                 timestamp=sc.timestamp)


def _render_with(raster_mod, sc):
    """import a fresh copy of the reference's gaussian_renderer bound to `raster_mod`"""
    sys.modules["diff_gaussian_rasterization_df"] = raster_mod
    sys.path.insert(0, CALLERS)
    try:
        for m in ("gaussian_renderer", "utils", "utils.sh_utils"):
            sys.modules.pop(m, None)
        gr = importlib.import_module("gaussian_renderer")
    finally:
        sys.path.remove(CALLERS)
    pc = StubModel(sc)
    pipe = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    out = gr.render(_camera(sc), pc, pipe, sc.bg.cuda(), near=sc.cam.min_depth, far=sc.cam.max_depth)
    go = synth.grad_outputs(sc)
    # train.py:148-153: colour loss + the hook tensor routed in as the gradient of "opticalflow"
    torch.autograd.backward([out["render"], out["opticalflow"]], [go["grad_color"].cuda(), go["grad_flow"].cuda()])
    res = {k: out[k].detach().cpu().numpy() for k in ("render", "depth", "opticalflow", "acc", "dominent_idxs", "radii")}
    res["visibility_filter"] = out["visibility_filter"].cpu().numpy()
    res["viewspace_grad"] = out["viewspace_points"].grad.cpu().numpy()
    res["l1points_grad"] = out["viewspace_l1points"].grad.cpu().numpy()
    res["xyz_grad"] = pc.inp["means3D"].grad.cpu().numpy()
    res["sh_grad"] = pc.inp["shs"].grad.cpu().numpy()
    for m in ("gaussian_renderer", "utils", "utils.sh_utils", "diff_gaussian_rasterization_df"):
        sys.modules.pop(m, None)
    return res


def test_unmodified_render_runs_on_the_dropin(built):
    if not os.path.exists(os.path.join(CALLERS, "gaussian_renderer", "__init__.py")):
        pytest.skip("reference caller not installed (python oracle/build_ref.py)")
    import diff_gaussian_rasterization_df as ours_pkg
    sc = synth.make_config("C1d", pose="tilted")
    a = _render_with(ours_pkg, sc)
    assert a["render"].shape == (3, sc.cam.H, sc.cam.W) and a["radii"].shape == (sc.P,)
    assert a["visibility_filter"].sum() > 0 and np.isfinite(a["viewspace_grad"]).all()
    assert np.abs(a["l1points_grad"]).max() > 0        # the per-Gaussian back-projected error statistics
    ref = U.reference_module()
    if ref is None:
        return
    b = _render_with(ref, sc)
    for k in ("render", "depth", "opticalflow", "acc"):
        assert float(np.abs(a[k] - b[k]).max()) <= 1e-4, k
    for k in ("dominent_idxs", "radii", "visibility_filter"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("viewspace_grad", "l1points_grad", "xyz_grad", "sh_grad"):
        assert U.rel_err(a[k], b[k], U.grad_floor(b[k])) <= 2e-3, k


def test_unmodified_render_with_fused_getters(built):
    """render() + FusedGetters (one kernel for all per-frame getters, SH handed over as the model's four
    tensors without torch.cat) == render() + PyTorch getters, values and gradients down to the parameters."""
    if not os.path.exists(os.path.join(CALLERS, "gaussian_renderer", "__init__.py")):
        pytest.skip("reference caller not installed (python oracle/build_ref.py)")
    import diff_gaussian_rasterization_df as ours_pkg
    from ex4dgs_b200.frontend import FusedGetters
    from ex4dgs_b200.rasterizer import SegmentedSH
    sc = synth.make_config("C1d", pose="tilted")

    def P(t):
        return t.detach().clone().cuda().requires_grad_(True)

    model = types.SimpleNamespace(
        _xyz=P(sc.xyz), _xyz_disp=P(sc.xyz_disp), _rotation=P(sc.rotation), _scaling=P(sc.scaling),
        _opacity=P(sc.opacity), _xyz_motion=P(sc.xyz_motion), _rotation_motion=P(sc.rotation_motion),
        _scaling_motion=P(sc.scaling_motion), _opacity_motion=P(sc.opacity_motion),
        _opacity_duration_center=P(sc.opacity_center), _opacity_duration_var=P(sc.opacity_var),
        _features_dc=P(sc.features[:, :1]), _features_rest=P(sc.features[:, 1:]),
        _features_dc_motion=P(sc.features_motion[:, :1]), _features_rest_motion=P(sc.features_motion[:, 1:]),
        duration=sc.duration, interval=sc.interval, time_shift=sc.time_shift, var_pad=sc.var_pad,
        kernel_size=sc.cam.kernel_size, active_sh_degree=sc.sh_degree, max_sh_degree=3)
    sys.modules["diff_gaussian_rasterization_df"] = ours_pkg
    sys.path.insert(0, CALLERS)
    try:
        for m in ("gaussian_renderer", "utils", "utils.sh_utils"):
            sys.modules.pop(m, None)
        gr = importlib.import_module("gaussian_renderer")
    finally:
        sys.path.remove(CALLERS)
    pipe = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    fg = FusedGetters(model)
    assert isinstance(fg.get_features(), SegmentedSH)
    out = gr.render(_camera(sc), fg, pipe, sc.bg.cuda(), near=sc.cam.min_depth, far=sc.cam.max_depth)
    go = synth.grad_outputs(sc)
    torch.autograd.backward([out["render"], out["opticalflow"]], [go["grad_color"].cuda(), go["grad_flow"].cuda()])
    ref = U.run_impl(ours_pkg, sc, kind="ours", grads=True, intermediates=False)
    assert float(np.abs(out["render"].detach().cpu().numpy() - ref["color"]).max()) <= 1e-4
    assert np.array_equal(out["radii"].cpu().numpy(), ref["radii"])
    # SH gradients arrive in the four parameter tensors, equal to the slices of the concatenated gradient
    Ns = sc.xyz.shape[0]
    gs = ref["grads"]["shs"]
    for got, want in ((model._features_dc.grad, gs[:Ns, :1]), (model._features_rest.grad, gs[:Ns, 1:]),
                      (model._features_dc_motion.grad, gs[Ns:, :1]), (model._features_rest_motion.grad, gs[Ns:, 1:])):
        assert got is not None and tuple(got.shape) == want.shape
        assert U.rel_err(got.cpu().numpy(), want, U.grad_floor(gs)) <= 2e-3
    assert model._xyz_motion.grad is not None and float(model._xyz_motion.grad.abs().max()) > 0
    for m in ("gaussian_renderer", "utils", "utils.sh_utils", "diff_gaussian_rasterization_df"):
        sys.modules.pop(m, None)
