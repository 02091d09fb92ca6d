"""Gradient parity AT SCALE: the backward compositing kernel against the compiled, unmodified reference at the sizes the
headline is quoted on (config 2: 500 k Gaussians, config 3: 2.0 M Gaussians, 1352x1014) and against the CPU oracle on a
one-tile torture scene whose list is 24 000 entries long (hundreds of wraps of the staging ring).

Tolerance: the north star's 1e-3 relative on every gradient tensor (|a - b| / max(|b|, floor), floor = 1 % of the tensor's
99th-percentile magnitude, tests/_util.py).  The reference's own backward is not bit-reproducible (float atomics in
arbitrary order: two runs of the reference on the same inputs differ by 1e-3 ... 7e-3 in single rotation / scale entries
at these sizes, a figure that itself changes from run to run), so each case runs the reference twice and the assertions
(_assert_within_reference_noise) are: all but 5 ppm of the entries of every tensor within 1e-3 of run 1 (or 4 x the
share by which run 2 misses run 1), none beyond 1e-2, relative L2 error <= 1e-4.  All figures are printed and
written to gpurun_out/parity_scale.json (DESIGN.md section 5 quotes them).
"""
import json
import os

import numpy as np
import pytest
import torch

from tests import _util as U
from tests.cases import make_one_tile_torture
from ex4dgs_b200 import synth

pytestmark = pytest.mark.gpu

GRAD_RTOL = 1e-3
OUT = os.environ.get("EX4DGS_PARITY_OUT") or \
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_scale.json")
_cache = {}


def _scene_and_reference(ref, cfg, grad_kind):
    """Scene + two runs of the reference (they do not depend on our cull mode): kept for the next parametrisation."""
    key = (cfg, grad_kind)
    if key not in _cache:
        _cache.clear()
        sc = synth.make_config(cfg, pose="tilted", dir_nonzero=(grad_kind == "all"), bg=torch.tensor([0.2, 0.7, 0.4]))
        r = U.run_impl(ref, sc, kind="ref", grads=True, grad_kind=grad_kind, intermediates=True)
        r2 = U.run_impl(ref, sc, kind="ref", grads=True, grad_kind=grad_kind, intermediates=False)
        _cache[key] = (sc, r, r2)
    return _cache[key]


def _record(name, rep):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    data = {}
    if os.path.exists(OUT):
        try:
            data = json.load(open(OUT))
        except Exception:
            data = {}
    data[name] = rep
    json.dump(data, open(OUT, "w"), indent=1)


def _stats(x, g, fl):
    d = np.abs(np.asarray(x, np.float64) - g) / np.maximum(np.abs(g), fl)
    g64 = np.asarray(g, np.float64)
    return dict(max_rel=float(d.max()), size=int(d.size), share_gt_1e3=float(np.mean(d > 1e-3)), share_gt_1e4=float(np.mean(d > 1e-4)),
                rel_l2=float(np.linalg.norm(np.asarray(x, np.float64) - g64) / max(np.linalg.norm(g64), 1e-300)))


def _grad_report(a, b, noise_of=None):
    rep = {}
    for k, g in b["grads"].items():
        fl = U.grad_floor(g)
        rep[k] = _stats(a["grads"][k], g, fl)
        rep[k]["floor"] = fl
        if noise_of is not None:
            n = _stats(noise_of["grads"][k], g, fl)
            rep[k].update(ref_vs_ref_max_rel=n["max_rel"], ref_vs_ref_share_gt_1e3=n["share_gt_1e3"], ref_vs_ref_rel_l2=n["rel_l2"])
    return rep


def _assert_within_reference_noise(name, rep):
    """The north star's 1e-3 on the gradients, read against a yardstick that does not reproduce itself to 1e-3 (the max
    over millions of entries of the reference-vs-reference deviation is 1e-3 ... 7e-3 and changes from run to run, ours
    moves with it).  Per tensor:
      * all but 5 ppm of the entries (or 4 x the reference's own share, or 3 entries) within 1e-3,
      * no entry beyond 1e-2,
      * relative L2 error <= 1e-4 (measured: ~1e-6, the reference's own run-to-run figure)."""
    for k, v in rep.items():
        print("%-30s %-10s max rel %.2e (ref run-to-run %.2e)  share > 1e-3: %.1e (ref %.1e)  rel L2 %.1e (ref %.1e)" %
              (name, k, v["max_rel"], v["ref_vs_ref_max_rel"], v["share_gt_1e3"], v["ref_vs_ref_share_gt_1e3"],
               v["rel_l2"], v["ref_vs_ref_rel_l2"]))
    for k, v in rep.items():
        assert v["max_rel"] <= 1e-2, (name, k, v)
        assert v["share_gt_1e3"] <= max(5e-6, 3.0 / v["size"], 4.0 * v["ref_vs_ref_share_gt_1e3"]), (name, k, v)
        assert v["rel_l2"] <= max(1e-4, 4.0 * v["ref_vs_ref_rel_l2"]), (name, k, v)


@pytest.mark.parametrize("cull", [0, 1], ids=["exact-lists", "tile-cull"])
@pytest.mark.parametrize("grad_kind", ["all", "color_flow"])
@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_gradients_against_compiled_reference_at_scale(built, cfg, cull, grad_kind):
    ref = U.reference_module()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    mod = U.ours_module()
    sc, r, r2 = _scene_and_reference(ref, cfg, grad_kind)
    old = mod.get_default_flags()
    mod.set_default_flags(bool(cull))
    try:
        ours = U.run_impl(mod, sc, kind="ours", grads=True, grad_kind=grad_kind, intermediates=not cull)
    finally:
        mod.set_default_flags(bool(old))
    assert ours["inexact_thresholds"] == 0
    assert np.array_equal(ours["radii"], r["radii"])
    assert np.array_equal(ours["idxs"], r["idxs"])
    for k in ("color", "depth", "acc", "flow"):
        assert np.array_equal(ours[k].view(np.uint32), r[k].view(np.uint32)), k + " not bit-identical"
    if not cull:
        assert np.array_equal(ours["inter"]["point_list"], r["inter"]["point_list"])
        assert np.array_equal(ours["inter"]["n_contrib"], r["inter"]["n_contrib"])
    rep = _grad_report(ours, r, noise_of=r2)
    name = "%s/%s/%s" % (cfg, "tile-cull" if cull else "exact-lists", grad_kind)
    _record(name, rep)
    _assert_within_reference_noise(name, rep)


@pytest.mark.parametrize("grad_kind", ["all", "color_flow"])
@pytest.mark.parametrize("cull", [0, 1], ids=["exact-lists", "tile-cull"])
def test_one_tile_torture(built, cull, grad_kind):
    """24 000 faint sub-pixel splats on one 16x16 tile: list of 24 000 entries, every pixel receives thousands of
    contributions; 375 sub-batches of 64 through the 4-deep mbarrier ring of the backward.  Integers / lists / image
    against the CPU oracle and, bit for bit, the compiled reference.  Gradients: every Gaussian here sums thousands of
    terms of mixed sign, in three different orders (sequential on the CPU, atomics in the reference, butterfly + warp
    reductions here); the two independent implementations of the SAME order-free mathematics - oracle and reference -
    differ from each other by up to 1.5e-3 in single rotation entries, which is printed as the yardstick."""
    mod = U.ours_module()
    sc = make_one_tile_torture()
    old = mod.get_default_flags()
    mod.set_default_flags(bool(cull))
    try:
        ours = U.run_impl(mod, sc, kind="ours", grads=True, grad_kind=grad_kind)
    finally:
        mod.set_default_flags(bool(old))
    orc = U.run_impl(U.oracle_module(), sc, dev="cpu", kind="oracle", grads=True, grad_kind=grad_kind)
    assert np.array_equal(ours["radii"], orc["radii"])
    assert np.array_equal(ours["idxs"], orc["idxs"])
    if not cull:
        assert np.array_equal(ours["inter"]["n_contrib"], orc["inter"]["n_contrib"])
        assert np.array_equal(ours["inter"]["point_list"], orc["inter"]["point_list"])
        assert np.array_equal(ours["inter"]["ranges"], orc["inter"]["ranges"])
    assert int(ours["inter"]["n_contrib"].max()) >= 20000
    for k in ("color", "depth", "acc", "flow"):
        assert float(np.abs(ours[k] - orc[k]).max()) <= 1e-4 * max(1.0, float(np.abs(orc[k]).max())), k
    name = "torture/%s/%s" % ("tile-cull" if cull else "exact-lists", grad_kind)

    def check(tag, rep):
        # max over 96 000 ill-conditioned sums is a noisy statistic (it moves between two runs of the same code): the
        # bounds are 1e-2 on the max, 1e-4 on the share of entries beyond 1e-3 and 1e-4 on the relative L2 error
        for k, v in rep.items():
            print("%-28s %-10s vs %-9s max rel %.2e  share > 1e-3: %.1e  rel L2 %.1e" % (name, k, tag, v["max_rel"], v["share_gt_1e3"], v["rel_l2"]))
            assert v["max_rel"] <= 1e-2 and v["share_gt_1e3"] <= 1e-4 and v["rel_l2"] <= 1e-4, (name, tag, k, v)

    rep = _grad_report(ours, orc)
    ref = U.reference_module()
    if ref is not None:
        r = U.run_impl(ref, sc, kind="ref", grads=True, grad_kind=grad_kind, intermediates=False)
        for k in ("color", "depth", "acc", "flow"):
            assert np.array_equal(ours[k].view(np.uint32), r[k].view(np.uint32)), k + " not bit-identical to the reference"
        rep_ref = _grad_report(ours, r)
        rep_or = _grad_report(orc, r)
        for k in rep_ref:
            rep_ref[k]["oracle_vs_ref_max_rel"] = rep_or[k]["max_rel"]
        _record(name + "/vs_reference", rep_ref)
        check("reference", rep_ref)
        check("(oracle vs reference)", rep_or)
    _record(name, rep)
    check("oracle", rep)
