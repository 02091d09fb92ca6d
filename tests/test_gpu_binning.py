"""The forward never waits for the instance count before it launches the binning: the binning buffer is sized from a
capacity hint and the count stays on the device (csrc/api.cu, csrc/binning.cu).  These tests drive every sizing path -
no hint (first frame: wait, then size exactly), hint too small (truncated first attempt, exact second run), hint far
too large - and the opt-in forward without any host wait (EX4DGS_FLAG_NO_HOST_WAIT) inside a CUDA graph.

Reference behaviour: rasterizer_impl.cu:293-336 (one blocking read of num_rendered, buffers sized exactly)."""
import numpy as np
import pytest
import torch

import ex4dgs_b200
from ex4dgs_b200 import synth
from oracle import getters_oracle as GO
from tests import _util as U

pytestmark = pytest.mark.gpu

INT_KEYS = ("radii", "idxs")
IMG_KEYS = ("color", "depth", "acc", "flow")


def _same(a, b, lists=True):
    for k in INT_KEYS:
        assert np.array_equal(a[k], b[k]), k
    for k in IMG_KEYS:
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    if lists:
        assert a["inter"]["R"] == b["inter"]["R"]
        for k in ("point_list", "tile_sorted", "ranges", "n_contrib"):
            assert np.array_equal(a["inter"][k], b["inter"][k]), k


@pytest.mark.parametrize("cull", [0, 1], ids=["exact-lists", "tile-cull"])
def test_every_sizing_path_gives_the_same_lists(built, cull):
    """500 k Gaussians, ~1.4 M instances: no hint / hint of 1000 instances (capacity 66 786 < R: overflow and second
    run) / hint of 3 R.  Lists, ranges, image: bit-identical; the exact-lists run is also checked against the oracle's
    order on a small case by the golden tests."""
    mod = U.ours_module()
    sc = synth.make_config("C2", pose="tilted")
    old = mod.get_default_flags()
    mod.set_default_flags(bool(cull))
    try:
        ex4dgs_b200.set_capacity_hint(0)
        first = U.run_impl(mod, sc, kind="ours", grads=True)
        R = first["inter"]["R"]
        assert R > 200000
        ex4dgs_b200.set_capacity_hint(1000)
        small = U.run_impl(mod, sc, kind="ours", grads=True)
        ex4dgs_b200.set_capacity_hint(3 * R)
        large = U.run_impl(mod, sc, kind="ours", grads=True)
        steady = U.run_impl(mod, sc, kind="ours", grads=True)       # hint left by the previous frame
    finally:
        mod.set_default_flags(bool(old))
        ex4dgs_b200.set_capacity_hint(0)
    _same(small, first)
    _same(large, first)
    _same(steady, first)
    # sortedness of the list that came out of the overflow path: (tile, depth bits, id) ascending
    it = small["inter"]
    d = it["depths"][it["point_list"]].view(np.uint32).astype(np.int64)
    key = (it["tile_sorted"].astype(np.int64) << 32) | d
    assert (np.diff(key) >= 0).all()
    assert (it["tile_sorted"] != 0xFFFF).all()           # rejected instances (dump tile) never reach the sorted list
    assert (np.diff(it["point_list"].astype(np.int64))[np.diff(key) == 0] > 0).all()


def test_nothing_visible_and_tiny_scenes(built):
    """R = 0 (everything behind the camera) and a scene of three Gaussians through every sizing path."""
    mod = U.ours_module()
    sc = synth.make_scene(P_static=300, P_dynamic=0, W=64, H=48, sigma_px=2.0)
    sc.xyz[:, 2] = -sc.xyz[:, 2].abs() - 1.0           # behind the camera
    for hint in (0, 1000):
        ex4dgs_b200.set_capacity_hint(hint)
        r = U.run_impl(mod, sc, kind="ours", grads=True)
        assert r["inter"]["R"] == 0 and (r["radii"] == 0).all() and (r["idxs"] == -1).all()
        assert all(np.all(v == 0) for v in r["grads"].values())
    sc3 = synth.make_scene(P_static=3, P_dynamic=0, W=64, H=48, sigma_px=4.0)
    ex4dgs_b200.set_capacity_hint(0)
    a = U.run_impl(mod, sc3, kind="ours", grads=True)
    b = U.run_impl(mod, sc3, kind="ours", grads=True)
    _same(a, b)
    ex4dgs_b200.set_capacity_hint(0)


def _rasterizer(sc, dev):
    return ex4dgs_b200.GaussianRasterizer(U.settings_for(ex4dgs_b200, sc, dev))


def _frame(rast, t, dev):
    means2D = torch.zeros(t["means3D"].shape[0], 3, device=dev, requires_grad=True)
    return rast(means3D=t["means3D"], means2D=means2D, dir3D=t["dir3D"], opacities=t["opacities"],
                shs=t["shs"], scales=t["scales"], rotations=t["rotations"])


NAMES = ("means3D", "dir3D", "opacities", "shs", "scales", "rotations")


def _leaves(inp, dev):
    return {k: inp[k].detach().clone().to(dev).requires_grad_(True) for k in NAMES}


def _eager(rast, t, dev, gc, gf):
    color, radii, depth, flow, acc, idxs = _frame(rast, t, dev)
    grads = torch.autograd.grad([color, flow], [t[k] for k in NAMES], [gc, gf])
    return [x.detach().clone() for x in (color, depth, acc, flow)] + [radii.clone(), idxs.clone()], [g.clone() for g in grads], color


def test_forward_in_a_cuda_graph(built):
    """EX4DGS_FLAG_NO_HOST_WAIT: the forward contains no host wait (a cudaEventSynchronize / cudaStreamSynchronize inside
    a stream capture fails it), so it can be captured into a CUDA graph and replayed on new inputs."""
    dev = torch.device("cuda", 0)
    sc = synth.make_config("C1d", pose="tilted")
    inp = GO.flat_inputs(sc)
    rast = _rasterizer(sc, dev)
    t_ref = {k: inp[k].detach().clone().to(dev) for k in NAMES}
    ex4dgs_b200.set_capacity_hint(0)
    ref = [x.clone() for x in _frame(rast, t_ref, dev)]          # also leaves the capacity hint for the captured frame
    ex4dgs_b200.set_host_wait(False)
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            t = {k: inp[k].detach().clone().to(dev) for k in NAMES}
            _frame(rast, t, dev)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            out = _frame(rast, t, dev)
        graph.replay()
        torch.cuda.synchronize()
        for a, b in zip(out, ref):
            assert torch.equal(a, b)
        with torch.no_grad():
            for d in (t, t_ref):
                d["means3D"] += 0.01
                d["opacities"] -= 0.1
        graph.replay()
        torch.cuda.synchronize()
        got = [x.clone() for x in out]
        ex4dgs_b200.set_host_wait(True)
        ref2 = _frame(rast, t_ref, dev)
        for a, b in zip(got, ref2):
            assert torch.equal(a, b)
        assert not torch.equal(got[0], ref[0])
    finally:
        ex4dgs_b200.set_host_wait(True)
        ex4dgs_b200.set_capacity_hint(0)


def test_forward_and_backward_in_a_cuda_graph(built):
    """Forward + backward of one frame captured together and replayed on new inputs."""
    dev = torch.device("cuda", 0)
    sc = synth.make_config("C1d", pose="tilted")
    inp = GO.flat_inputs(sc)
    go = synth.grad_outputs(sc)
    gc, gf = go["grad_color"].to(dev), go["grad_flow"].to(dev)
    rast = _rasterizer(sc, dev)
    t_ref = _leaves(inp, dev)
    ex4dgs_b200.set_capacity_hint(0)
    ref_out, ref_grads, color = _eager(rast, t_ref, dev, gc, gf)
    R = int(color.grad_fn.num_rendered)
    ex4dgs_b200.set_host_wait(False)
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):        # leaves and warm-up on the capture stream (their grad accumulators must not live on the default stream)
            t = _leaves(inp, dev)
            _eager(rast, t, dev, gc, gf)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            color, radii, depth, flow, acc, idxs = _frame(rast, t, dev)
            grads = torch.autograd.grad([color, flow], [t[k] for k in NAMES], [gc, gf])
        assert int(color.grad_fn.num_rendered) >= R          # the capacity stands in for the count
        graph.replay()
        torch.cuda.synchronize()
        assert not ex4dgs_b200.frame_overflowed(color)
        for a, b in zip((color, depth, acc, flow), ref_out[:4]):
            assert torch.equal(a, b)
        assert torch.equal(radii, ref_out[4]) and torch.equal(idxs, ref_out[5])
        for k, a, b in zip(NAMES, grads, ref_grads):
            assert U.rel_err(a.cpu().numpy(), b.cpu().numpy(), U.grad_floor(b.cpu().numpy())) <= 1e-3, k
        with torch.no_grad():
            for d in (t, t_ref):
                d["means3D"] += 0.01
                d["opacities"] -= 0.1
        graph.replay()
        torch.cuda.synchronize()
        got = [x.clone() for x in (color, depth, acc, flow)]
        got_grads = [g.clone() for g in grads]
        ex4dgs_b200.set_host_wait(True)
        ref_out2, ref_grads2, _ = _eager(rast, t_ref, dev, gc, gf)
        for a, b in zip(got, ref_out2[:4]):
            assert torch.equal(a, b)
        for k, a, b in zip(NAMES, got_grads, ref_grads2):
            assert U.rel_err(a.cpu().numpy(), b.cpu().numpy(), U.grad_floor(b.cpu().numpy())) <= 1e-3, k
    finally:
        ex4dgs_b200.set_host_wait(True)
        ex4dgs_b200.set_capacity_hint(0)


def test_forward_without_host_wait_flags_a_truncated_frame(built):
    dev = torch.device("cuda", 0)
    sc = synth.make_config("C2", pose="tilted")
    inp = GO.flat_inputs(sc)
    t = _leaves(inp, dev)
    rast = _rasterizer(sc, dev)
    try:
        ex4dgs_b200.set_capacity_hint(0)
        ex4dgs_b200.set_host_wait(False)
        with pytest.raises(RuntimeError, match="capacity hint"):
            _frame(rast, t, dev)
        ex4dgs_b200.set_capacity_hint(1000)                  # capacity 66 786, the frame has ~1.4 M instances
        color = _frame(rast, t, dev)[0]
        assert ex4dgs_b200.frame_overflowed(color)
        ex4dgs_b200.set_capacity_hint(4000000)
        color = _frame(rast, t, dev)[0]
        assert not ex4dgs_b200.frame_overflowed(color)
        ex4dgs_b200.set_host_wait(True)
        exact = _frame(rast, t, dev)[0]
        assert torch.equal(color, exact)
    finally:
        ex4dgs_b200.set_host_wait(True)
        ex4dgs_b200.set_capacity_hint(0)


def test_more_than_65535_tiles_against_oracle(built):
    """4112 x 4112 pixels = 257 x 257 = 66 049 tiles: tile ids no longer fit 16-bit keys (and 0xFFFF is a real tile), so the
    duplicate kernel emits 32-bit keys and the tile sort runs three 8-bit passes over them (the reference sorts 64-bit
    keys whatever the image size, rasterizer_impl.cu:303-323).  Against the CPU oracle in both modes of the exact tile
    culling: radii, tiles_touched, point_list, ranges and n_contrib bit-equal; images within 5e-4 (25-pixel splats: long
    per-pixel sums through glibc's expf on the oracle side) with the arg-max id differing in at most 1e-6 of the 16.9 M
    pixels (near-ties of two weights); gradients at 1e-3.  Against the compiled reference, when it is installed: the
    image and the ids bit-identical."""
    mod = U.ours_module()
    sc = synth.make_scene(4000, 0, 4112, 4112, sigma_px=25.0, seed=synth.SEED + 77)
    orc = U.run_impl(U.oracle_module(), sc, dev="cpu", kind="oracle", grad_kind="all")
    assert (orc["inter"]["ranges"][65535:] != 0).any()
    ref = U.reference_module()
    r = U.run_impl(ref, sc, kind="ref", grads=False, intermediates=False) if ref is not None else None
    old = mod.get_default_flags()
    try:
        for cull in (0, 1):
            mod.set_default_flags(bool(cull))
            ex4dgs_b200.set_capacity_hint(0)
            ours = U.run_impl(mod, sc, kind="ours", grad_kind="all")
            it, io = ours["inter"], orc["inter"]
            assert it["tile_sorted"].dtype == np.uint32 and int(it["tile_sorted"].max()) > 65535
            assert ours.get("inexact_thresholds", 0) == 0
            assert np.array_equal(ours["radii"], orc["radii"])
            assert float((ours["idxs"] != orc["idxs"]).mean()) <= 1e-6
            if not cull:
                for k in ("tiles_touched", "point_list", "ranges", "n_contrib"):
                    assert it["R"] == io["R"] and np.array_equal(it[k], io[k]), k
            for k in ("color", "depth", "acc", "flow"):
                assert float(np.abs(ours[k] - orc[k]).max()) <= 5e-4 * max(1.0, float(np.abs(orc[k]).max())), k
            U.assert_grads_close(ours["grads"], orc["grads"], 1e-3,
                                 rerun=lambda: U.run_impl(mod, sc, kind="ours", grad_kind="all", intermediates=False)["grads"])
            if r is not None:
                assert np.array_equal(ours["color"].view(np.uint32), r["color"].view(np.uint32)), "image not bit-identical"
                assert np.array_equal(ours["idxs"], r["idxs"]) and np.array_equal(ours["radii"], r["radii"])
            steady = U.run_impl(mod, sc, kind="ours", grad_kind="all")       # capacity from the previous frame
            _same(steady, ours)
    finally:
        mod.set_default_flags(bool(old))
        ex4dgs_b200.set_capacity_hint(0)
