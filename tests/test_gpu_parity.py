"""GPU parity tests: the CUDA path (through the C ABI / public Python API) against
  (a) the CPU oracle on seeded scenes the oracle finishes in seconds,
  (b) the committed golden vectors produced by the unmodified reference (tests/golden/),
  (c) the compiled reference itself when oracle/_ref travelled to the box,
  (d) size-independent properties at the full BASELINE.json sizes.

Tolerances are the north star's: tile/key indexing bit-exact, RGB <= 1e-4 abs, gradients <= 1e-3
relative (with an absolute floor of 1 % of the tensor's 99th-percentile magnitude: the reference's
own float-atomic backward is not reproducible below that).
"""
import os

import numpy as np
import pytest
import torch

from tests import _util as U
from tests.cases import GOLDEN_CASES, make_case
from ex4dgs_b200 import synth
from oracle import getters_oracle as GO  # noqa: E402

pytestmark = pytest.mark.gpu

RGB_TOL = 1e-4
GRAD_RTOL = 1e-3      # the north star's figure (the backward decides pair membership exactly like the forward, render_bwd.cu)
GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(params=[0, 1], ids=["exact-lists", "tile-cull"])
def cull(request):
    """Run the parity cases in both modes of EX4DGS_FLAG_TILE_CULL: with the flag clear the tile
    lists must equal the reference's bit for bit; with it set only the user-visible outputs and the
    gradients are compared (the internal lists are shorter by construction)."""
    mod = U.ours_module()
    old = mod.get_default_flags()
    mod.set_default_flags(bool(request.param))
    yield request.param
    mod.set_default_flags(bool(old))


def _check_ints(a, b, cull=0):
    # every Gaussian's skip threshold is the exact crossing of the alpha >= 1/255 test (preprocess.cu alpha_threshold)
    assert a.get("inexact_thresholds", 0) == 0
    assert np.array_equal(a["radii"], b["radii"])
    assert np.array_equal(a["idxs"], b["idxs"])
    if cull:
        return
    ia, ib = a["inter"], b["inter"]
    assert ia["R"] == ib["R"]
    assert np.array_equal(ia["tiles_touched"], ib["tiles_touched"])
    assert np.array_equal(ia["point_list"], ib["point_list"])
    assert np.array_equal(ia["ranges"], ib["ranges"])
    assert np.array_equal(ia["n_contrib"], ib["n_contrib"])
    assert np.array_equal(a["idxs"], b["idxs"])
    vis = b["radii"] > 0
    assert np.array_equal(ia["depths"][vis].view(np.uint32), ib["depths"][vis].view(np.uint32))
    assert np.array_equal(ia["means2D"][vis].view(np.uint32), ib["means2D"][vis].view(np.uint32))


def _check_floats(a, b, grads=True, rerun=None, rerun_ref=None):
    """rerun / rerun_ref: callables that render the frame again (U.run_impl) - see U.assert_grads_close."""
    for k in ("color", "depth", "acc", "flow"):
        scale = max(1.0, float(np.abs(b[k]).max()))
        assert float(np.abs(a[k] - b[k]).max()) <= RGB_TOL * scale, k
    if grads:
        U.assert_grads_close(a["grads"], b["grads"], GRAD_RTOL, rerun=(lambda: rerun()["grads"]) if rerun else None,
                             rerun_ref=(lambda: rerun_ref()["grads"]) if rerun_ref else None)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_against_cpu_oracle(built, name, cull):
    sc, kw = make_case(name)
    ours = U.run_impl(U.ours_module(), sc, kind="ours", **kw)
    orc = U.run_impl(U.oracle_module(), sc, dev="cpu", kind="oracle", **kw)
    _check_ints(ours, orc, cull)
    _check_floats(ours, orc, rerun=lambda: U.run_impl(U.ours_module(), sc, kind="ours", **kw))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_against_reference_golden(built, name, cull):
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet")
    g = np.load(path)
    sc, kw = make_case(name)
    ours = U.run_impl(U.ours_module(), sc, kind="ours", **kw)
    ref = {k: g[k] for k in ("color", "radii", "depth", "flow", "acc", "idxs")}
    ref["inter"] = {k[6:]: g[k] for k in g.files if k.startswith("inter_")}
    ref["inter"]["R"] = int(ref["inter"]["R"])
    ref["grads"] = {k[5:]: g[k] for k in g.files if k.startswith("grad_")}
    _check_ints(ours, ref, cull)
    _check_floats(ours, ref, rerun=lambda: U.run_impl(U.ours_module(), sc, kind="ours", **kw))
    if cull:
        return
    # the reference's 64-bit keys are (tile << 32 | depth bits) of our lists
    keys = (ours["inter"]["tile_sorted"].astype(np.uint64) << np.uint64(32)) | \
        ours["inter"]["depths"][ours["inter"]["point_list"]].view(np.uint32).astype(np.uint64)
    assert np.array_equal(keys, ref["inter"]["point_list_keys"])


@pytest.mark.parametrize("cfg,kw", [("C1", {}), ("C1d", dict(pose="tilted", dir_nonzero=True))])
def test_config1_against_oracle(built, cfg, kw, cull):
    sc = synth.make_config(cfg, **kw)
    ours = U.run_impl(U.ours_module(), sc, kind="ours", grad_kind="all")
    orc = U.run_impl(U.oracle_module(), sc, dev="cpu", kind="oracle", grad_kind="all")
    _check_ints(ours, orc, cull)
    _check_floats(ours, orc, rerun=lambda: U.run_impl(U.ours_module(), sc, kind="ours", grad_kind="all"))


@pytest.mark.parametrize("cfg", ["C1d", "C2"])
def test_live_against_compiled_reference(built, cfg, cull):
    ref = U.reference_module()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    sc = synth.make_config(cfg, pose="tilted", dir_nonzero=True, bg=torch.tensor([0.2, 0.7, 0.4]))
    grads = cfg != "C2"                       # config 2 is forward-only (BASELINE.json)
    ours = U.run_impl(U.ours_module(), sc, kind="ours", grads=True, grad_kind="all")
    r = U.run_impl(ref, sc, kind="ref", grads=True, grad_kind="all")
    _check_ints(ours, r, cull)
    assert np.array_equal(ours["color"].view(np.uint32), r["color"].view(np.uint32)), "image not bit-identical"
    _check_floats(ours, r, grads=grads, rerun=lambda: U.run_impl(U.ours_module(), sc, kind="ours", grads=True, grad_kind="all"),
                  rerun_ref=lambda: U.run_impl(ref, sc, kind="ref", grads=True, grad_kind="all"))


def test_full_size_properties(built):
    """Config 3 (2.0M Gaussians, 1352x1014): properties that need no second implementation."""
    sc = synth.make_config("C3")
    mod = U.ours_module()
    old_flags = mod.get_default_flags()
    mod.set_default_flags(False)
    a = U.run_impl(mod, sc, kind="ours")
    ia = a["inter"]
    R, P = ia["R"], sc.P
    W, H = sc.cam.W, sc.cam.H
    # scan / duplicate: R = sum of tiles_touched; every list entry is a visible Gaussian
    assert R == int(ia["tiles_touched"].astype(np.int64).sum())
    assert (a["radii"][ia["point_list"]] > 0).all()
    # ranges partition [0, R) in tile order and the tile of every entry matches its range
    rg = ia["ranges"].astype(np.int64)
    nonempty = rg[:, 1] > rg[:, 0]
    assert (rg[~nonempty] == 0).all()
    starts, ends = rg[nonempty, 0], rg[nonempty, 1]
    assert starts[0] == 0 and ends[-1] == R and (starts[1:] == ends[:-1]).all()
    tiles_of_entry = np.repeat(np.nonzero(nonempty)[0], (ends - starts))
    assert np.array_equal(tiles_of_entry, ia["tile_sorted"].astype(np.int64))
    # sortedness: within a tile, (depth bits, id) ascending  == the reference's stable 64-bit key sort
    d = ia["depths"][ia["point_list"]].view(np.uint32).astype(np.int64)
    key = (ia["tile_sorted"].astype(np.int64) << 32) | d
    assert (np.diff(key) >= 0).all()
    same = np.diff(key) == 0
    assert (np.diff(ia["point_list"].astype(np.int64))[same] > 0).all()
    # compositing invariants
    assert (a["acc"] >= 0).all() and (a["acc"] <= 1.0 + 1e-5).all()
    assert np.allclose(ia["final_T"].reshape(H, W) + a["acc"][0], 1.0, atol=2e-4)
    assert ((a["idxs"] == -1) == (a["acc"] == 0)).all()
    assert (ia["n_contrib"].reshape(H, W)[a["acc"][0] == 0] == 0).all()
    # determinism of the forward (bit-exact twice) and of the integer outputs
    b = U.run_impl(mod, sc, kind="ours", grads=False, intermediates=False)
    for k in ("color", "depth", "acc", "flow", "idxs", "radii"):
        assert np.array_equal(a[k], b[k]), k
    # exact-output tile culling: same outputs and gradients (to atomic-order noise), shorter lists
    mod.set_default_flags(True)
    try:
        c = U.run_impl(mod, sc, kind="ours")
    finally:
        mod.set_default_flags(bool(old_flags))
    for k in ("color", "depth", "acc", "flow", "idxs", "radii"):
        assert np.array_equal(a[k], c[k]), k
    kept = int((c["inter"]["ranges"][:, 1].astype(np.int64) - c["inter"]["ranges"][:, 0]).sum())
    # the scan counts the ellipse's bounding-box rectangle (tight_rect), the exact test then moves rejected
    # instances to the dump tile behind every range
    Rc = c["inter"]["R"]
    assert kept <= Rc < R and Rc == int(c["inter"]["tiles_touched"].astype(np.int64).sum())
    print("C3 instances: reference rectangles %d, bounding-box rectangles %d, after the exact tile test %d" % (R, Rc, kept))
    mod.set_default_flags(False)
    try:
        a2 = U.run_impl(mod, sc, kind="ours", intermediates=False)      # exact-list mode again: the run-to-run yardstick
    finally:
        mod.set_default_flags(bool(old_flags))
    for k, g in a["grads"].items():
        e, noise = U.rel_err(c["grads"][k], g, U.grad_floor(g)), U.rel_err(a2["grads"][k], g, U.grad_floor(g))
        print("%-10s cull-vs-exact max rel %.2e   exact-vs-exact %.2e" % (k, e, noise))
        assert e <= 1e-3 + noise, k
    # linearity of the backward in the upstream gradient (checksum-of-checksums style property)
    assert np.isfinite(a["grads"]["means3D"]).all()


def test_edge_cases(built, cull):
    mod = U.ours_module()
    dev = "cuda"
    # P == 0: outputs keep the reference's fill values (rasterize_points.cu:73-90)
    sc = synth.make_scene(8, 0, 64, 48)
    rs = U.settings_for(mod, sc, dev)
    e = torch.zeros(0, 3, device=dev)
    out = mod.GaussianRasterizer(rs)(means3D=e, means2D=e, dir3D=e, opacities=torch.zeros(0, 1, device=dev),
                                     shs=torch.zeros(0, 16, 3, device=dev), scales=e, rotations=torch.zeros(0, 4, device=dev))
    color, radii, depth, flow, acc, idxs = out
    assert color.shape == (3, 48, 64) and float(color.abs().max()) == 0.0 and int(idxs.max()) == -1 and radii.numel() == 0
    # everything culled (behind the camera): R == 0, image = background, depth = max_depth
    sc2 = synth.make_scene(50, 0, 64, 48, bg=torch.tensor([0.3, 0.6, 0.9]))
    sc2.xyz[:, 2] = -5.0
    r = U.run_impl(mod, sc2, kind="ours")
    assert r["inter"]["R"] == 0 and (r["radii"] == 0).all()
    assert np.allclose(r["color"], np.array([0.3, 0.6, 0.9], np.float32)[:, None, None])
    assert np.allclose(r["depth"], sc2.cam.max_depth) and (r["idxs"] == -1).all()
    assert all(float(np.abs(g).max()) == 0.0 for g in r["grads"].values())
    orc = U.run_impl(U.oracle_module(), sc2, dev="cpu", kind="oracle")
    assert np.array_equal(r["color"], orc["color"])
    # one huge splat covering the whole image (cooperative duplicate path, long single-entry lists)
    sc3 = synth.make_scene(3, 0, 200, 120, sigma_px=120.0, seed=7)
    a = U.run_impl(mod, sc3, kind="ours")
    b = U.run_impl(U.oracle_module(), sc3, dev="cpu", kind="oracle")
    _check_ints(a, b, cull)
    _check_floats(a, b, rerun=lambda: U.run_impl(mod, sc3, kind="ours"))


def test_mark_visible(built):
    mod = U.ours_module()
    from oracle import oracle as orc
    sc = synth.make_config("C1", pose="tilted")
    inp = GO.flat_inputs(sc)
    rs = U.settings_for(mod, sc, "cuda")
    vis = mod.GaussianRasterizer(rs).markVisible(inp["means3D"].cuda()).cpu().numpy()
    ref = orc.mark_visible(inp["means3D"].numpy(), sc.cam.viewmatrix.numpy(), sc.cam.projmatrix.numpy(),
                           sc.cam.min_depth, sc.cam.max_depth)
    assert np.array_equal(vis, ref)
    assert 0 < vis.sum() < vis.size


def test_non_default_stream_and_reentrancy(built, cull):
    """The library launches on torch's current stream (the reference uses the legacy default stream)."""
    mod = U.ours_module()
    sc, kw = make_case("gold_base")
    base = U.run_impl(mod, sc, kind="ours", **kw)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        other = U.run_impl(mod, sc, kind="ours", **kw)
    s.synchronize()
    assert np.array_equal(base["color"], other["color"])
    assert np.array_equal(base["inter"]["point_list"], other["inter"]["point_list"])


@pytest.mark.parametrize("variant", ["deg0", "deg2", "scale_mod", "kernel_size", "tight_planes", "odd_size"])
def test_setting_variants_against_oracle(built, variant, cull):
    """Less common GaussianRasterizationSettings values, each against the CPU oracle."""
    kw = dict(P_static=900, P_dynamic=300, W=112, H=80, sigma_px=3.0, seed=synth.SEED + 40, pose="tilted")
    scale_modifier = 1.0
    if variant == "odd_size":
        kw.update(W=97, H=53)
    if variant == "tight_planes":
        kw.update(min_depth=3.0, max_depth=25.0)
    if variant == "kernel_size":
        kw.update(kernel_size=0.35)
    sc = synth.make_scene(**kw)
    if variant == "deg0":
        sc.sh_degree = 0
    if variant == "deg2":
        sc.sh_degree = 2
    if variant == "scale_mod":
        scale_modifier = 1.7

    def run(mod, dev, kind):
        if scale_modifier == 1.0:
            return U.run_impl(mod, sc, dev=dev, kind=kind, grad_kind="all")
        orig = U.settings_for

        def patched(m, s_, d, subpixel=None, debug=False):
            rs = orig(m, s_, d, subpixel, debug)
            return rs._replace(scale_modifier=scale_modifier)
        U.settings_for = patched
        try:
            return U.run_impl(mod, sc, dev=dev, kind=kind, grad_kind="all")
        finally:
            U.settings_for = orig
    ours = run(U.ours_module(), "cuda", "ours")
    orc = run(U.oracle_module(), "cpu", "oracle")
    _check_ints(ours, orc, cull)
    _check_floats(ours, orc, rerun=lambda: run(U.ours_module(), "cuda", "ours"))


def test_debug_flag_and_error_paths(built):
    """debug=True synchronises after every stage (auxiliary.h:296-303) and must give the same result;
    bad SH layouts are rejected by the C ABI with a message."""
    mod = U.ours_module()
    sc, kw = make_case("gold_base")
    base = U.run_impl(mod, sc, kind="ours", **kw)
    orig = U.settings_for
    U.settings_for = lambda m, s_, d, subpixel=None, debug=False: orig(m, s_, d, subpixel, True)
    try:
        dbg = U.run_impl(mod, sc, kind="ours", **kw)
    finally:
        U.settings_for = orig
    assert np.array_equal(base["color"], dbg["color"])
    inp = GO.flat_inputs(sc)
    rs = U.settings_for(mod, sc, "cuda")
    P = inp["means3D"].shape[0]
    with pytest.raises(RuntimeError, match="needs 16 coefficients"):
        mod.GaussianRasterizer(rs)(means3D=inp["means3D"].cuda(), means2D=torch.zeros(P, 3).cuda(), dir3D=inp["dir3D"].cuda(),
                                   opacities=inp["opacities"].cuda(), shs=inp["shs"][:, :9].contiguous().cuda(),
                                   scales=inp["scales"].cuda(), rotations=inp["rotations"].cuda())


@pytest.mark.parametrize("seed", [101, 202, 303, 404])
def test_bit_exact_stress_against_reference(built, seed):
    """Forward of 4 x 500k random Gaussians (config-2 size, tilted camera, subpixel offsets, random
    background): every integer output, the tile lists and the image must equal the compiled
    reference's bit for bit (exact-list mode)."""
    ref = U.reference_module()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    mod = U.ours_module()
    old = mod.get_default_flags()
    mod.set_default_flags(False)
    try:
        g = torch.Generator().manual_seed(seed)
        sc = synth.make_config("C2", seed=seed, pose="tilted", dir_nonzero=True, bg=torch.rand(3, generator=g))
        sub = torch.rand(sc.cam.H, sc.cam.W, 2, generator=g) - 0.5
        a = U.run_impl(mod, sc, kind="ours", grads=True, subpixel=sub)
        b = U.run_impl(ref, sc, kind="ref", grads=True, subpixel=sub)
    finally:
        mod.set_default_flags(bool(old))
    _check_ints(a, b, 0)
    for k in ("color", "depth", "acc", "flow"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    vis = b["radii"] > 0
    assert np.array_equal(a["inter"]["conic_opacity"][vis].view(np.uint32), b["inter"]["conic_opacity"][vis].view(np.uint32))
    assert np.array_equal(a["inter"]["rgb"][vis].view(np.uint32), b["inter"]["rgb"][vis].view(np.uint32))


def test_flow_free_forward_equals_general_kernel(built, cull):
    """dir3D all +-0 (what gaussian_renderer/__init__.py:66 always passes) selects the compositing kernel
    without flow accumulators; a single non-zero component on a VISIBLE Gaussian selects the general one.
    Every non-flow output and internal list must be bit-identical between the two, the flow image must be
    exactly +0 in the first case, and a non-zero dir3D on a culled Gaussian must not switch kernels' results."""
    mod = U.ours_module()
    dev = "cuda"
    sc = synth.make_config("C1")
    inp = {k: v.to(dev) for k, v in GO.flat_inputs(sc).items()}
    P = inp["means3D"].shape[0]
    rs = U.settings_for(mod, sc, dev)

    def run(dir3D):
        with torch.no_grad():
            return [t.clone() for t in mod.GaussianRasterizer(rs)(
                means3D=inp["means3D"], means2D=torch.zeros(P, 3, device=dev), dir3D=dir3D, opacities=inp["opacities"],
                shs=inp["shs"], scales=inp["scales"], rotations=inp["rotations"])]

    z = torch.zeros(P, 3, device=dev)
    base = run(z)
    radii = base[1]
    assert float(base[3].abs().max()) == 0.0 and not torch.signbit(base[3]).any()
    vis = int(torch.nonzero(radii > 0)[0])
    hid = int(torch.nonzero(radii == 0)[0])
    neg = z.clone()
    neg[::3] = -0.0                                             # negative zeros still count as "no flow"
    a = run(neg)
    one = z.clone()
    one[vis, 1] = 0.75                                          # general kernel
    b = run(one)
    hidden = z.clone()
    hidden[hid] = torch.tensor([1.0, float("nan"), 2.0])        # culled Gaussian: contributes nothing
    c = run(hidden)
    for other in (a, b, c):
        for i in (0, 1, 2, 4, 5):                               # color, radii, depth, acc, idxs
            assert torch.equal(base[i], other[i]), i
    assert float(a[3].abs().max()) == 0.0 and float(c[3].abs().max()) == 0.0
    assert float(b[3].abs().max()) > 0.0
    # and the general kernel agrees with the CPU oracle on that flow image
    sc2 = synth.make_config("C1")
    ref = U.oracle_module()
    rso = U.settings_for(ref, sc2, "cpu")
    inc = GO.flat_inputs(sc2)
    with torch.no_grad():
        fo = ref.GaussianRasterizer(rso)(means3D=inc["means3D"], means2D=torch.zeros(P, 3), dir3D=one.cpu(),
                                         opacities=inc["opacities"], shs=inc["shs"], scales=inc["scales"],
                                         rotations=inc["rotations"])[3]
    assert float((fo - b[3].cpu()).abs().max()) <= 1e-5


@pytest.mark.parametrize("split", ["C1d", "all-static", "all-dynamic", "odd"])
def test_segmented_sh_equals_concatenated(built, cull, split):
    """EX4DGS_FLAG_SH_SEGMENTED: the model's four SH tensors read in place == the [P,16,3] torch.cat of
    get_features() (scene/c_gaussian_model.py:337-353): bit-identical forward, gradients equal to the
    slices of dL_dsh; static/dynamic boundaries inside a 128-Gaussian block, empty segments, SH degree 1."""
    from ex4dgs_b200.rasterizer import SegmentedSH
    mod = U.ours_module()
    dev = "cuda"
    sc = synth.make_config("C1d", pose="tilted")
    if split == "odd":
        sc.sh_degree = 1
    inp = {k: v.to(dev) for k, v in GO.flat_inputs(sc).items()}
    P = inp["means3D"].shape[0]
    Ns = {"C1d": sc.xyz.shape[0], "all-static": P, "all-dynamic": 0, "odd": 4099}[split]
    rs = U.settings_for(mod, sc, dev)
    go = {k: v.to(dev) for k, v in synth.grad_outputs(sc).items()}

    def run(shs):
        t = {k: inp[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations")}
        m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
        d3 = torch.zeros(P, 3, device=dev, requires_grad=True)
        out = mod.GaussianRasterizer(rs)(means3D=t["means3D"], means2D=m2, dir3D=d3, opacities=t["opacities"], shs=shs,
                                         scales=t["scales"], rotations=t["rotations"])
        torch.autograd.backward([out[0], out[2], out[3], out[4]],
                                [go["grad_color"], go["grad_depth"], go["grad_flow"], go["grad_acc"]])
        return out, t, m2, d3

    full = inp["shs"].clone().requires_grad_(True)
    a_out, a_t, a_m2, a_d3 = run(full)
    parts = [inp["shs"][:Ns, :1].clone().contiguous().requires_grad_(True), inp["shs"][:Ns, 1:].clone().contiguous().requires_grad_(True),
             inp["shs"][Ns:, :1].clone().contiguous().requires_grad_(True), inp["shs"][Ns:, 1:].clone().contiguous().requires_grad_(True)]
    seg = SegmentedSH(*parts)
    assert tuple(seg.shape) == (P, 16, 3) and torch.equal(seg.cat(), inp["shs"])
    b_out, b_t, b_m2, b_d3 = run(seg)
    for x, y in zip(a_out, b_out):
        assert torch.equal(x, y)
    gs = full.grad.cpu().numpy()
    fl = U.grad_floor(gs)
    want = [gs[:Ns, :1], gs[:Ns, 1:], gs[Ns:, :1], gs[Ns:, 1:]]
    for g, w in zip(parts, want):
        assert g.grad is not None and tuple(g.grad.shape) == w.shape
        assert U.rel_err(g.grad.cpu().numpy(), w, fl) <= 2e-3
    for k in a_t:
        ga = a_t[k].grad.cpu().numpy()
        assert U.rel_err(b_t[k].grad.cpu().numpy(), ga, U.grad_floor(ga)) <= 2e-3, k
    assert U.rel_err(b_m2.grad.cpu().numpy(), a_m2.grad.cpu().numpy(), U.grad_floor(a_m2.grad.cpu().numpy())) <= 2e-3
    # argument checks
    with pytest.raises(ValueError):
        SegmentedSH(parts[0], parts[0], parts[2], parts[3])
    with pytest.raises(RuntimeError):
        mod.GaussianRasterizer(rs)(means3D=inp["means3D"][:-1], means2D=torch.zeros(P - 1, 3, device=dev), dir3D=torch.zeros(P - 1, 3, device=dev),
                                   opacities=inp["opacities"][:-1], shs=seg, scales=inp["scales"][:-1], rotations=inp["rotations"][:-1])


@pytest.mark.parametrize("used", ["color", "color+flow", "color+acc", "color+depth", "flow"])
def test_absent_upstream_gradients_equal_zero_gradients(built, cull, used):
    """Outputs the loss does not use reach backward as None (set_materialize_grads(False)), cross the C ABI
    as NULL and select the compositing backward without depth / acc terms; the reference gets zero tensors
    for them (__init__.py:110-178).  Every gradient must equal the one computed with explicit zeros (same
    kernel arithmetic, only the exactly-zero terms are dropped), and must agree with the CPU oracle."""
    mod = U.ours_module()
    dev = "cuda"
    sc = synth.make_config("C1d", pose="tilted")
    inp = {k: v.to(dev) for k, v in GO.flat_inputs(sc).items()}
    P = inp["means3D"].shape[0]
    rs = U.settings_for(mod, sc, dev)
    go = {k: v.to(dev) for k, v in synth.grad_outputs(sc).items()}
    g = torch.Generator().manual_seed(5)
    go["grad_depth"] = (torch.randn(1, sc.cam.H, sc.cam.W, generator=g) * 1e-3).to(dev)
    go["grad_acc"] = (torch.randn(1, sc.cam.H, sc.cam.W, generator=g) * 1e-3).to(dev)
    names = ("grad_color", "grad_depth", "grad_flow", "grad_acc")
    on = {"grad_color": "color" in used, "grad_depth": "depth" in used, "grad_flow": "flow" in used, "grad_acc": "acc" in used}

    def run(explicit_zeros):
        t = {k: inp[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "shs")}
        m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
        d3 = torch.zeros(P, 3, device=dev, requires_grad=True)
        out = mod.GaussianRasterizer(rs)(means3D=t["means3D"], means2D=m2, dir3D=d3, opacities=t["opacities"], shs=t["shs"],
                                         scales=t["scales"], rotations=t["rotations"])
        outs = dict(grad_color=out[0], grad_depth=out[2], grad_flow=out[3], grad_acc=out[4])
        if explicit_zeros:
            torch.autograd.backward([outs[n] for n in names], [go[n] if on[n] else torch.zeros_like(go[n]) for n in names])
        else:
            torch.autograd.backward([outs[n] for n in names if on[n]], [go[n] for n in names if on[n]])
        res = {k: v.grad.cpu().numpy() for k, v in t.items()}
        res["means2D"] = m2.grad.cpu().numpy()
        res["dir3D"] = d3.grad.cpu().numpy()
        return res

    a, b = run(True), run(False)
    for k in a:
        assert U.rel_err(b[k], a[k], U.grad_floor(a[k])) <= 2e-3, k
    if used == "color":
        assert float(np.abs(b["dir3D"]).max()) == 0.0           # no flow gradient: dL_ddir3D is exactly zero


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_tile_cull_is_output_exact_on_adversarial_scenes(built, seed):
    """The exact-output claim of EX4DGS_FLAG_TILE_CULL (bounding-box rectangles + exact tile test + dump tile) on scenes
    built to stress it: needle-like splats (anisotropy up to 300:1) at every orientation, opacities spread around the
    1/255 visibility limit, splats hundreds of pixels wide, centres far outside the image, subpixel offsets of up to
    +-3 pixels.  Every user-visible output must be bit-identical with the flag clear (reference rectangles and lists)
    and set; gradients equal to summation-order noise."""
    mod = U.ours_module()
    g = torch.Generator().manual_seed(seed)
    sc = synth.make_config("C1", seed=seed, pose="tilted", sigma_px=3.0)
    n = sc.xyz.shape[0]
    # elongate: one axis x up to 300, another / up to 10; a tenth of the splats becomes huge, a tenth tiny
    stretch = torch.exp(torch.empty(n, 3).uniform_(-2.3, 5.7, generator=g) * (torch.rand(n, 3, generator=g) < 0.4))
    size = torch.ones(n, 1)
    r = torch.rand(n, generator=g)
    size[r < 0.1] = 30.0
    size[r > 0.9] = 0.05
    sc.scaling = (sc.scaling + torch.log(stretch) + torch.log(size)).contiguous()
    # opacities: a third around the 1/255 limit (logit(1/255) = -5.54), the rest anywhere
    lim = torch.rand(n, generator=g) < 0.33
    sc.opacity = torch.where(lim[:, None], -5.54 + 0.5 * torch.randn(n, 1, generator=g), 3.0 * torch.randn(n, 1, generator=g)).contiguous()
    # a fifth of the centres pushed towards / beyond the frustum border
    far = torch.rand(n, generator=g) < 0.2
    sc.xyz = torch.where(far[:, None], sc.xyz * torch.tensor([1.35, 1.35, 1.0]), sc.xyz).contiguous()
    sub = (torch.rand(sc.cam.H, sc.cam.W, 2, generator=g) - 0.5) * (6.0 if seed != 13 else 0.0)
    old = mod.get_default_flags()
    try:
        mod.set_default_flags(False)
        a = U.run_impl(mod, sc, kind="ours", grads=True, subpixel=sub, grad_kind="all")
        a2 = U.run_impl(mod, sc, kind="ours", grads=True, subpixel=sub, grad_kind="all", intermediates=False)
        mod.set_default_flags(True)
        b = U.run_impl(mod, sc, kind="ours", grads=True, subpixel=sub, grad_kind="all")
    finally:
        mod.set_default_flags(bool(old))
    assert int((a["radii"] > 0).sum()) > n // 4
    for k in ("color", "depth", "acc", "flow"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    for k in ("radii", "idxs"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["inter"]["n_contrib"] > 0, b["inter"]["n_contrib"] > 0)
    assert b["inter"]["R"] < a["inter"]["R"]
    # gradients: the float reductions run in a different order (other list lengths, other sub-batches).  On these scenes
    # the covariance chain of needle-like splats amplifies that rounding noise without bound (scale / rotation gradients
    # of single splats differ by O(1) between two runs of the SAME mode), so the yardstick is the same mode run twice
    # and the statistic is the share of entries off by more than 1e-3.
    def off_share(x, ref):
        fl = U.grad_floor(ref)
        d = np.abs(np.asarray(x, np.float64) - ref) / np.maximum(np.abs(ref), fl)
        return float(np.mean(d > 1e-3))
    for k, ga in a["grads"].items():
        noise, err = off_share(a2["grads"][k], ga), off_share(b["grads"][k], ga)
        print("%-10s share of entries off by > 1e-3: cull-vs-exact %.2e   exact-vs-exact %.2e" % (k, err, noise))
        assert err <= max(1e-4, 3.0 * noise), (k, err, noise)


def test_nine_coefficient_sh_rows_against_oracle(built, cull):
    """shs given as [P,9,3] (degree-2 rows, M = 9): the forward reads rows of M coefficients and the backward takes the
    generic per-Gaussian kernel (preprocess_bwd_kernel; the staged kernel handles the 16-coefficient layout).  Both share
    their derivative blocks with the staged kernel (bwd_* functions of csrc/preprocess.cu)."""
    mod, orc = U.ours_module(), U.oracle_module()
    sc = synth.make_scene(900, 300, 112, 80, sigma_px=3.0, seed=synth.SEED + 41, pose="tilted")
    sc.sh_degree = 2
    go = synth.grad_outputs(sc)

    def run(m, dev):
        inp = {k: v.to(dev) for k, v in GO.flat_inputs(sc).items()}
        inp["shs"] = inp["shs"][:, :9].contiguous()
        t = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
        P = t["means3D"].shape[0]
        m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
        out = m.GaussianRasterizer(U.settings_for(m, sc, dev))(means3D=t["means3D"], means2D=m2, dir3D=t["dir3D"], opacities=t["opacities"],
                                                             shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
        torch.autograd.backward([out[0], out[2], out[3], out[4]], [go[k].to(dev) for k in ("grad_color", "grad_depth", "grad_flow", "grad_acc")])
        res = {k: v.grad.detach().cpu().numpy() for k, v in t.items()}
        res["means2D"] = m2.grad.detach().cpu().numpy()
        return out[0].detach().cpu().numpy(), out[1].cpu().numpy(), res

    ca, ra, ga = run(mod, "cuda")
    cb, rb, gb = run(orc, "cpu")
    assert np.array_equal(ra, rb)
    assert float(np.abs(ca - cb).max()) <= RGB_TOL
    assert ga["shs"].shape == (1200, 9, 3)
    U.assert_grads_close(ga, gb, GRAD_RTOL, rerun=lambda: run(mod, "cuda")[2])
