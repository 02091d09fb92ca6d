"""GPU: the fused per-iteration statistics kernel and the regulariser kernel (csrc/stats.cu, through the C ABI
via ex4dgs_b200/stats.py) against the fixture made by the reference's own code and against the CPU oracle."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import stats_oracle as SO
from oracle.make_stats_golden import fresh_state, make_inputs
from ex4dgs_b200 import stats

pytestmark = pytest.mark.gpu
FIX = os.path.join(os.path.dirname(__file__), "golden", "stats_fixture.npz")
EXACT = ("max_radii2D", "min_radii2D", "denom", "error_denom", "xyz_error_min_timestamp", "motion_max_radii2D",
         "motion_min_radii2D", "motion_denom", "motion_error_denom", "motion_xyz_error_min_timestamp")


def _model(Ns, Nd, dev):
    m = SimpleNamespace(**{k: v.to(dev) for k, v in fresh_state(Ns, Nd).items()})
    m._xyz = torch.zeros(Ns, 3, device=dev)
    return m


def _compare(m, want, tag=""):
    for k in SO.ALL_NAMES:
        got = getattr(m, k).cpu().numpy()
        w = np.asarray(want[k])
        assert got.shape == w.shape, k
        if k in EXACT:
            assert np.array_equal(got, w), tag + k
        else:       # accumulated float sums / IEEE quotients: sqrt(x*x + y*y) may be contracted differently
            assert np.allclose(got, w, rtol=2e-6, atol=1e-12), tag + k


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_stats_kernel_equals_reference_fixture(built, tag):
    fx = np.load(FIX)
    Ns, Nd, l1_accum, seed, steps = [int(v) for v in fx["%s_meta" % tag]]
    dev = torch.device("cuda")
    m = _model(Ns, Nd, dev)
    for it in make_inputs(Ns, Nd, seed, steps):
        stats.iteration_stats(m, it["radii"].to(dev), it["grad"].to(dev), it["err"].to(dev) if l1_accum else None,
                              it["timestamp"], densify=it["densify"])
    _compare(m, {k: fx["%s_%s" % (tag, k)] for k in SO.ALL_NAMES}, tag)


def test_stats_kernel_equals_oracle_large_and_is_sync_free(built):
    dev = torch.device("cuda")
    Ns, Nd = 70001, 29999
    m = _model(Ns, Nd, dev)
    state = fresh_state(Ns, Nd)
    side = torch.cuda.Stream()
    for it in make_inputs(Ns, Nd, 77, 3):
        SO.iteration_stats(state, Ns, it["radii"], it["grad"], it["err"], it["timestamp"], it["densify"])
        with torch.cuda.stream(side):                       # launches on the caller's current stream
            stats.iteration_stats(m, it["radii"].to(dev), it["grad"].to(dev), it["err"].to(dev), it["timestamp"],
                                  densify=it["densify"])
    side.synchronize()
    _compare(m, {k: state[k].numpy() for k in SO.ALL_NAMES})
    # NaN error gradients propagate like torch.clamp_min / comparisons do
    it = make_inputs(Ns, Nd, 78, 1)[0]
    it["err"][::97, 0] = float("nan")
    SO.iteration_stats(state, Ns, it["radii"], it["grad"], it["err"], 5.0, True)
    stats.iteration_stats(m, it["radii"].to(dev), it["grad"].to(dev), it["err"].to(dev), 5.0, densify=True)
    for k in SO.ALL_NAMES:
        a, b = getattr(m, k).cpu().numpy(), state[k].numpy()
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        assert np.allclose(a, b, rtol=2e-6, atol=1e-12, equal_nan=True), k


def test_stats_argument_checks(built):
    dev = torch.device("cuda")
    m = _model(8, 4, dev)
    r = torch.ones(12, dtype=torch.int32, device=dev)
    g = torch.zeros(12, 3, device=dev)
    with pytest.raises(RuntimeError):
        stats.iteration_stats(m, r.long(), g, None, 0.0)                        # radii must be the int32 output
    with pytest.raises(RuntimeError):
        stats.iteration_stats(m, r, g[:11], None, 0.0)
    with pytest.raises(RuntimeError):
        stats.iteration_stats(m, r[:5], g[:5], None, 0.0)                       # fewer radii than static Gaussians
    del m.motion_denom
    with pytest.raises(AttributeError):
        stats.iteration_stats(m, r, g, None, 0.0)
    stats.iteration_stats(_model(0, 0, dev), r[:0], g[:0], None, 0.0)           # empty model: nothing to do


def test_regularizers_equal_reference_fixture(built):
    fx = np.load(FIX)
    dev = torch.device("cuda")
    sr, mr = [float(v) for v in fx["reg_weights"]]
    d = torch.from_numpy(fx["reg_disp"]).to(dev).requires_grad_(True)
    mo = torch.from_numpy(fx["reg_motion"]).to(dev).requires_grad_(True)
    terms = stats.regularizers_(d, mo, sr, mr)                                  # .grad is None: gradients are written
    assert np.allclose(terms.cpu().numpy(), fx["reg_terms"], rtol=2e-6)
    assert np.allclose(d.grad.cpu().numpy(), fx["reg_gdisp"], rtol=1e-5, atol=1e-12)
    assert np.allclose(mo.grad.cpu().numpy(), fx["reg_gmotion"], rtol=1e-5, atol=1e-12)
    assert float(d.grad[5].abs().max()) == 0.0 and float(mo.grad[3, 4].abs().max()) == 0.0


@pytest.mark.parametrize("Ns,Nd,K", [(50000, 20000, 36), (1000, 0, 36), (33, 77, 2), (7, 5, 70)])
def test_regularizers_equal_oracle_and_accumulate(built, Ns, Nd, K):
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(Ns + Nd + K)
    disp = torch.randn(Ns, 3, generator=g) * 0.03
    motion = torch.randn(Nd, K, 3, generator=g) * 0.5
    sr, mr = 1e-4, 3e-4
    terms, gd, gm = SO.regularizers(disp, motion, sr, mr)
    d = disp.to(dev).requires_grad_(True)
    mo = motion.to(dev).requires_grad_(True)
    pre_d = torch.randn(Ns, 3, generator=g) * 1e-6
    pre_m = torch.randn(Nd, K, 3, generator=g) * 1e-7
    d.grad = pre_d.to(dev).clone()
    mo.grad = pre_m.to(dev).clone()
    scale = torch.tensor(0.5, device=dev)
    t = stats.regularizers_(d, mo, sr, mr, loss_grad=scale)                     # adds 0.5 * gradient to the existing .grad
    t2 = stats.regularizers_(d, mo, sr, mr, value_only=True)
    assert torch.equal(t, t2)                                                   # fixed-order reduction: bit-reproducible
    assert np.allclose(t.cpu().numpy(), terms.numpy(), rtol=3e-6, atol=1e-12)
    assert np.allclose(d.grad.cpu().numpy(), (pre_d.double() + 0.5 * gd).numpy(), rtol=2e-5, atol=1e-11)
    if Nd:
        assert np.allclose(mo.grad.cpu().numpy(), (pre_m.double() + 0.5 * gm).numpy(), rtol=2e-5, atol=1e-11)
    # weights of 0: terms off, gradients untouched
    before = d.grad.clone()
    t0 = stats.regularizers_(d, mo, 0.0, 0.0)
    assert float(t0.abs().sum()) == 0.0 and torch.equal(before, d.grad)


def test_stats_live_against_reference_class_on_gpu(built):
    """The reference's UNMODIFIED CGaussianModel methods (oracle/_ref/callers/scene/c_gaussian_model.py, installed by
    oracle/build_ref.py) and the verbatim train.py:196-212 block on CUDA tensors, against the fused kernel acting on a
    second instance of the same class: the model object is used as it is, attribute names and shapes included."""
    import bench
    cls = bench.load_reference_model_class()
    if cls is None:
        pytest.skip("oracle/_ref/callers not installed (no /root/reference at build time)")
    dev = torch.device("cuda")
    Ns, Nd = 4001, 1777

    def model():
        g = cls.__new__(cls)
        for k, v in fresh_state(Ns, Nd).items():
            setattr(g, k, v.to(dev))
        g._xyz = torch.zeros(Ns, 3, device=dev)
        return g

    ref, ours = model(), model()
    for it in make_inputs(Ns, Nd, 4321, 3):
        radii, grad, err = it["radii"].to(dev), it["grad"].to(dev), it["err"].to(dev)
        vp, ve = SimpleNamespace(grad=grad), SimpleNamespace(grad=err)
        gaussians, visibility_filter = ref, radii > 0
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
            gaussians.add_l1_ssim_stats(ve, static_vis_filter, dynamic_vis_filter, static_num, it["timestamp"])
        stats.iteration_stats(ours, radii, grad, err, it["timestamp"], densify=it["densify"])
    _compare(ours, {k: getattr(ref, k).cpu().numpy() for k in SO.ALL_NAMES}, "live ")
