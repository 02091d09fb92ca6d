"""ex4dgs_b200.densify (one-launch row gathers, csrc/compact.cu) against the reference's UNMODIFIED CGaussianModel
(oracle/_ref/callers/scene/c_gaussian_model.py:693-1072) on live CUDA tensors: two instances of the reference class hold
the same scene, the same optimizer state and the same statistics; one keeps its own `_prune_optimizer` / `prune_points` /
`cat_tensors_to_optimizer`, the other gets this repository's three methods bound on (`densify.install`).  Row copies
are exact, so everything is compared bit for bit: 15 parameters, 30 moments, step counters, 18 statistics tensors -
after a prune, after a concatenation, and after the reference's whole `densify_and_prune` (clone + split + prune with
its random samples drawn from the same seed)."""
import types

import pytest
import torch

from ex4dgs_b200 import densify, synth

pytestmark = pytest.mark.gpu

ARGS = types.SimpleNamespace(
    percent_dense=0.01, position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01,
    position_lr_max_steps=30000, dynamic_position_lr_init=0.00016, dynamic_position_lr_final=0.000016,
    dynamic_position_lr_delay_mult=0.01, dynamic_position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.05,
    scaling_lr=0.005, rotation_lr=0.00001, disp_lr=0.0001, feature_motion_lr=0.0025, rotation_motion_lr=0.001,
    opacity_motion_lr=0.05, opacity_motion_center_lr=0.001, opacity_motion_var_lr=0.0005)

PARAMS = tuple(densify._STATIC_ATTR.values()) + tuple(densify._DYNAMIC_ATTR.values())
STATS = densify.STATIC_STATS + densify.DYNAMIC_STATS


def _model(sc, cls, steps=2, seed=3):
    """Reference model + its own training_setup (torch.optim.RAdam over the 15 named groups), `steps` optimizer steps on
    seeded random gradients (non-trivial moments), seeded random statistics."""
    import bench
    m = bench.reference_model(sc, torch.device("cuda", 0), cls)
    m.spatial_lr_scale = 1.0
    m.keyframe_num = int(m._xyz_motion.shape[1])          # create_from_pcd / load_ply set it (c_gaussian_model.py:603)
    m.training_setup(ARGS)
    m.max_radii2D = torch.zeros(m._xyz.shape[0], device="cuda")
    m.min_radii2D = torch.ones(m._xyz.shape[0], device="cuda") * 1000
    m.motion_max_radii2D = torch.zeros(m._xyz_motion.shape[0], device="cuda")
    m.motion_min_radii2D = torch.ones(m._xyz_motion.shape[0], device="cuda") * 1000
    g = torch.Generator(device="cuda").manual_seed(seed)
    for _ in range(steps):
        for n in PARAMS:
            p = getattr(m, n)
            p.grad = torch.randn(p.shape, generator=g, device="cuda") * 1e-2
        m.optimizer.step()
        m.optimizer.zero_grad(set_to_none=True)
    for n in STATS:
        t = getattr(m, n)
        setattr(m, n, torch.rand(t.shape, generator=g, device="cuda"))
    return m


def _pair(sc, **kw):
    import bench
    cls = bench.load_reference_model_class()
    if cls is None:
        pytest.skip("oracle/_ref/callers not installed (no /root/reference at build time)")
    ref, ours = _model(sc, cls, **kw), _model(sc, cls, **kw)
    densify.install(ours)
    return ref, ours


def _assert_same(ref, ours):
    for n in PARAMS + STATS:
        a, b = getattr(ours, n), getattr(ref, n)
        assert a.shape == b.shape and a.dtype == b.dtype, (n, tuple(a.shape), tuple(b.shape))
        assert torch.equal(a, b), n
    for ga, gb in zip(ours.optimizer.param_groups, ref.optimizer.param_groups):
        assert ga["name"] == gb["name"] and ga["lr"] == gb["lr"]
        pa, pb = ga["params"][0], gb["params"][0]
        assert isinstance(pa, torch.nn.Parameter) and pa.requires_grad and pa.is_leaf and torch.equal(pa, pb), ga["name"]
        sa, sb = ours.optimizer.state.get(pa), ref.optimizer.state.get(pb)
        assert (sa is None) == (sb is None), ga["name"]
        if sa is not None:
            assert float(sa["step"]) == float(sb["step"])
            for k in ("exp_avg", "exp_avg_sq"):
                assert sa[k].shape == pa.shape and torch.equal(sa[k], sb[k]), (ga["name"], k)
    assert len(ours.optimizer.state) == len(ref.optimizer.state)
    for n, attr in list(densify._STATIC_ATTR.items()) + list(densify._DYNAMIC_ATTR.items()):
        group = [g for g in ours.optimizer.param_groups if g["name"] == n][0]
        if getattr(ours, attr).shape[0] or n in densify._STATIC_ATTR:
            assert group["params"][0] is getattr(ours, attr), n      # the model trains what the optimizer steps


def test_prune_points_matches_the_reference_class(built):
    sc = synth.make_scene(5000, 3000, 64, 48, seed=11)
    ref, ours = _pair(sc)
    g = torch.Generator(device="cuda").manual_seed(1)
    ms = torch.rand(5000, generator=g, device="cuda") < 0.3
    md = torch.rand(3000, generator=g, device="cuda") < 0.6
    ref.prune_points(ms, md)
    ours.prune_points(ms, md)
    assert ours._xyz.shape[0] == int((~ms).sum()) and ours._xyz_motion.shape[0] == int((~md).sum())
    _assert_same(ref, ours)
    # a second round on the compacted model, then a step of the optimizer on the new parameters
    ms2 = torch.rand(ours._xyz.shape[0], generator=g, device="cuda") < 0.5
    md2 = torch.zeros(ours._xyz_motion.shape[0], dtype=torch.bool, device="cuda")          # nothing removed
    ref.prune_points(ms2, md2)
    ours.prune_points(ms2, md2)
    _assert_same(ref, ours)
    for m in (ref, ours):
        for n in PARAMS:
            p = getattr(m, n)
            p.grad = torch.full_like(p, 1e-3)
        m.optimizer.step()
    _assert_same(ref, ours)


def test_mask_of_the_wrong_length_is_refused(built):
    """torch refuses x[mask] when the mask's length differs from the tensor's; the gather kernel does not range-check its
    index, so the drop-in must refuse it on the host too (a statistics tensor that was not extended after a densification)."""
    sc = synth.make_scene(600, 400, 64, 48, seed=15)
    ref, ours = _pair(sc)
    for m in (ref, ours):
        m.xyz_error_min = m.xyz_error_min[:-3]
    ms, md = torch.zeros(600, dtype=torch.bool, device="cuda"), torch.zeros(400, dtype=torch.bool, device="cuda")
    with pytest.raises(IndexError):
        ref.prune_points(ms, md)
    with pytest.raises(IndexError):
        ours.prune_points(ms, md)
    assert ours._xyz.shape[0] == 600 and ours.optimizer.param_groups[0]["params"][0] is ours._xyz      # nothing was touched
    with pytest.raises(IndexError):
        ours.prune_points(ms[:-1], md)


def test_prune_everything_and_nothing(built):
    sc = synth.make_scene(700, 500, 64, 48, seed=12)
    ref, ours = _pair(sc)
    none_s, none_d = torch.zeros(700, dtype=torch.bool, device="cuda"), torch.zeros(500, dtype=torch.bool, device="cuda")
    ref.prune_points(none_s, none_d)
    ours.prune_points(none_s, none_d)
    _assert_same(ref, ours)
    ref.prune_points(none_s, ~none_d)          # every dynamic Gaussian goes
    ours.prune_points(none_s, ~none_d)
    assert ours._xyz_motion.shape == (0,) + tuple(ref._xyz_motion.shape[1:])
    _assert_same(ref, ours)
    # no dynamic Gaussians left: the reference passes an empty float tensor as the dynamic mask (c_gaussian_model.py:1049)
    ms = torch.arange(700, device="cuda") % 3 == 0
    ref.prune_points(ms, torch.empty(0).cuda())
    ours.prune_points(ms, torch.empty(0).cuda())
    _assert_same(ref, ours)


def test_cat_tensors_to_optimizer_matches_the_reference_class(built):
    sc = synth.make_scene(3000, 2000, 64, 48, seed=13)
    ref, ours = _pair(sc)
    g = torch.Generator(device="cuda").manual_seed(2)
    ext = {}
    for name, attr in list(densify._STATIC_ATTR.items()) + list(densify._DYNAMIC_ATTR.items()):
        p = getattr(ref, attr)
        ext[name] = torch.randn((401 if name in densify._STATIC_ATTR else 77,) + tuple(p.shape[1:]), generator=g, device="cuda")
    a = ref.cat_tensors_to_optimizer(ext)
    b = ours.cat_tensors_to_optimizer(ext)
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k
        assert torch.equal(ours.optimizer.state[b[k]]["exp_avg"][-5:], torch.zeros_like(b[k][-5:])), k
    # a subset of the groups (densification_postfix_onlystatic, c_gaussian_model.py:846-872) through the reference's own caller
    sub = [torch.randn((9,) + tuple(getattr(ref, attr).shape[1:]), generator=g, device="cuda") for attr in densify._STATIC_ATTR.values()]
    for m, d in ((ref, a), (ours, b)):
        for name, attr in list(densify._STATIC_ATTR.items()) + list(densify._DYNAMIC_ATTR.items()):
            setattr(m, attr, d[name])
        m.densification_postfix_onlystatic(*sub)
    for n in PARAMS:
        assert torch.equal(getattr(ours, n), getattr(ref, n)), n
    assert ours._xyz.shape[0] == 3000 + 401 + 9


@pytest.mark.parametrize("dynamic", [True, False], ids=["static+dynamic", "static-only"])
def test_whole_densify_and_prune_of_the_reference_runs_on_top(built, dynamic):
    """CGaussianModel.densify_and_prune (c_gaussian_model.py:1019-1072: densify_and_clone, densify_and_split, prune_points)
    UNCHANGED, once with the class's own tensor surgery and once with ours bound underneath, random samples from the same
    seed: identical models."""
    sc = synth.make_scene(6000, 2500 if dynamic else 0, 64, 48, seed=14)
    ref, ours = _pair(sc)
    for m in (ref, ours):
        # statistics in the ranges training produces, so that every branch selects some Gaussians
        g = torch.Generator(device="cuda").manual_seed(9)
        m.xyz_gradient_accum = torch.rand(m._xyz.shape[0], 1, generator=g, device="cuda") * 4e-4
        m.denom = torch.ones(m._xyz.shape[0], 1, device="cuda")
        m.motion_xyz_gradient_accum = torch.rand(m._xyz_motion.shape[0], 1, generator=g, device="cuda") * 4e-4
        m.motion_denom = torch.ones(m._xyz_motion.shape[0], 1, device="cuda")
        m.max_radii2D = torch.rand(m._xyz.shape[0], generator=g, device="cuda") * 40
        m.motion_max_radii2D = torch.rand(m._xyz_motion.shape[0], generator=g, device="cuda") * 40
        m.xyz_error_accum *= 0.05
        m.motion_xyz_error_mean *= 0.05
    n0 = ours._xyz.shape[0]
    for m in (ref, ours):
        torch.manual_seed(1234)
        torch.cuda.manual_seed(1234)
        m.densify_and_prune(0.0002, 0.0001, 0.005, 0.005, 5.0, 20, 20)
    assert ours._xyz.shape[0] != n0
    _assert_same(ref, ours)


def test_gather_rows_function(built):
    g = torch.Generator(device="cuda").manual_seed(4)
    ts = [torch.randn(1000, generator=g, device="cuda"), torch.randn(1000, 3, generator=g, device="cuda"),
          torch.randn(1000, 15, 3, generator=g, device="cuda"), torch.randn(1000, 36, 4, generator=g, device="cuda"),
          torch.randint(0, 100, (1000, 2), generator=g, device="cuda", dtype=torch.int32),
          torch.randn(1001, 36, 3, generator=g, device="cuda")[1:]]                     # 4-byte aligned only
    idx = torch.randint(0, 1000, (2500,), generator=g, device="cuda")
    outs = densify.gather_rows(ts, idx)
    for t, o in zip(ts, outs):
        assert torch.equal(o, t[idx])
    assert densify.gather_rows(ts, idx[:0])[2].shape == (0, 15, 3)
    with pytest.raises(RuntimeError):
        densify.gather_rows([torch.zeros(10, 3)], torch.zeros(2, dtype=torch.int64))     # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        densify.gather_rows([torch.zeros(10, 3, dtype=torch.uint8, device="cuda")], idx[:2] % 10)   # 3-byte rows


@pytest.mark.parametrize("n,e", [(1000, 7), (1001, 7), (1000, 0), (0, 5), (4096, 4096)])
def test_concatenation_jobs_run_flat(built, n, e):
    """Jobs without an index are two contiguous copies (or a copy and a zero fill): 128-bit flat path when the split is a
    multiple of four words, scalar otherwise, 1-3 tail words behind the split."""
    g = torch.Generator(device="cuda").manual_seed(n + e)
    jobs = densify._Jobs()
    want, outs = [], []
    for shape in ((3,), (15, 3), (36, 4), ()):
        a = torch.randn((n,) + shape, generator=g, device="cuda")
        b = torch.randn((e,) + shape, generator=g, device="cuda")
        outs.append(jobs.add(a, n + e, n_a=n, b=b))
        want.append(torch.cat((a, b)))
        outs.append(jobs.add(a, n + e, n_a=n))                      # zero rows appended (the moments)
        want.append(torch.cat((a, torch.zeros_like(b))))
    jobs.launch()
    for o, w in zip(outs, want):
        assert o.shape == w.shape and torch.equal(o, w)
