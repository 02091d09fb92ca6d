"""CPU tests of the oracle (oracle/cpu_raster.c): pinned against the golden vectors produced by the
unmodified reference on a B200 (tests/golden/*.npz, generator: oracle/make_golden.py), plus
properties the reference algorithm implies."""
import os

import numpy as np
import pytest
import torch

from tests import _util as U
from tests.cases import GOLDEN_CASES, make_case
from ex4dgs_b200 import synth
from oracle import getters_oracle as GO  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_golden(built, name):
    """Integer / index outputs bit-exact, floats within the north-star tolerances."""
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet")
    g = np.load(path)
    sc, kw = make_case(name)
    o = U.run_impl(U.oracle_module(), sc, dev="cpu", kind="oracle", **kw)
    st = o["inter"]
    assert np.array_equal(o["radii"], g["radii"])
    assert st["R"] == int(g["inter_R"])
    assert np.array_equal(st["tiles_touched"], g["inter_tiles_touched"])
    assert np.array_equal(st["point_list_keys"], g["inter_point_list_keys"])      # 64-bit keys: tile | depth bits
    assert np.array_equal(st["point_list"], g["inter_point_list"])
    assert np.array_equal(st["ranges"], g["inter_ranges"])
    assert np.array_equal(st["n_contrib"], g["inter_n_contrib"])
    assert np.array_equal(o["idxs"], g["idxs"])
    vis = g["radii"] > 0
    assert np.array_equal(st["depths"][vis].view(np.uint32), g["inter_depths"][vis].view(np.uint32))
    assert np.array_equal(st["means2D"][vis].view(np.uint32), g["inter_means2D"][vis].view(np.uint32))
    assert U.rel_err(st["conic_opacity"][vis], g["inter_conic_opacity"][vis], 1e-6) <= 1e-5
    if "colors" not in o["grads"]:
        assert float(np.abs(st["rgb"][vis] - g["inter_rgb"][vis]).max()) <= 1e-5
        assert np.array_equal(st["clamped"][vis], g["inter_clamped"][vis])
    for k in ("color", "depth", "acc", "flow"):
        scale = max(1.0, float(np.abs(g[k]).max()))
        assert float(np.abs(o[k] - g[k]).max()) <= 1e-4 * scale, k
    assert float(np.abs(st["final_T"] - g["inter_final_T"]).max()) <= 1e-5
    for k, v in o["grads"].items():
        ref = g["grad_" + k]
        assert U.rel_err(v, ref, U.grad_floor(ref)) <= 2e-3, k


def test_oracle_empty_and_all_culled(built):
    from oracle import oracle as orc
    sc = synth.make_scene(20, 0, 40, 24, bg=torch.tensor([0.25, 0.5, 0.75]))
    cam = sc.cam
    kw = dict(bg=sc.bg.numpy(), W=cam.W, H=cam.H, viewmatrix=cam.viewmatrix.numpy(), projmatrix=cam.projmatrix.numpy(),
              campos=cam.campos.numpy(), tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, kernel_size=0.1, subpixel_offset=None,
              min_depth=0.2, max_depth=100.0)
    o = orc.Oracle()
    out = o.forward(means3D=np.zeros((0, 3), np.float32), dir3D=None, opacities=None, shs=None, **kw)
    assert out["R"] == 0 and float(np.abs(out["color"]).max()) == 0.0 and (out["idxs"] == -1).all()   # rasterize_points.cu:73-90
    inp = {k: v.numpy() for k, v in GO.flat_inputs(sc).items()}
    inp["means3D"][:, 2] = -3.0
    out = o.forward(means3D=inp["means3D"], dir3D=inp["dir3D"], opacities=inp["opacities"], shs=inp["shs"],
                    scales=inp["scales"], rotations=inp["rotations"], **kw)
    assert out["R"] == 0 and (out["radii"] == 0).all()
    assert np.allclose(out["color"], sc.bg.numpy()[:, None, None]) and np.allclose(out["depth"], 100.0)


def test_oracle_backward_is_linear_in_upstream_gradients(built):
    """dL/dtheta(a*g1 + b*g2) = a*dL/dtheta(g1) + b*dL/dtheta(g2) - except the cumulative dL_dacc
    path, which is still linear (A.3-Q4 multiplies by T, not by the gradient)."""
    from oracle import oracle as orc
    sc, _ = make_case("gold_tilted_all")
    inp = {k: v.numpy() for k, v in GO.flat_inputs(sc).items()}
    cam = sc.cam
    o = orc.Oracle()
    o.forward(bg=sc.bg.numpy(), W=cam.W, H=cam.H, means3D=inp["means3D"], dir3D=inp["dir3D"], opacities=inp["opacities"],
              shs=inp["shs"], scales=inp["scales"], rotations=inp["rotations"], viewmatrix=cam.viewmatrix.numpy(),
              projmatrix=cam.projmatrix.numpy(), campos=cam.campos.numpy(), tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
              kernel_size=cam.kernel_size, subpixel_offset=None, min_depth=cam.min_depth, max_depth=cam.max_depth)
    rng = np.random.default_rng(0)
    H, W = cam.H, cam.W

    def rnd():
        return [rng.standard_normal((3, H, W)).astype(np.float32), rng.standard_normal((1, H, W)).astype(np.float32),
                rng.standard_normal((3, H, W)).astype(np.float32), rng.standard_normal((1, H, W)).astype(np.float32)]
    g1, g2 = rnd(), rnd()
    a, b = 0.5, -2.0
    r1, r2 = o.backward(*g1), o.backward(*g2)
    r3 = o.backward(*[a * x + b * y for x, y in zip(g1, g2)])
    for k in r1:
        comb = a * r1[k] + b * r2[k]
        assert U.rel_err(r3[k], comb, U.grad_floor(comb)) <= 1e-3, k


def test_oracle_colour_gradients_are_true_derivatives(built):
    """Where the reference's backward IS the true derivative (SH / colour and dir3D paths) the
    hand-written backward must agree with central finite differences of the oracle's forward."""
    from oracle import oracle as orc
    sc = synth.make_scene(60, 20, 48, 32, sigma_px=4.0, seed=3, dir_nonzero=True)
    inp = {k: v.numpy().astype(np.float32) for k, v in GO.flat_inputs(sc).items()}
    cam = sc.cam
    rng = np.random.default_rng(1)
    gc = rng.standard_normal((3, cam.H, cam.W)).astype(np.float32)
    gf = rng.standard_normal((3, cam.H, cam.W)).astype(np.float32)
    zero1 = np.zeros((1, cam.H, cam.W), np.float32)

    def fwd(shs, dir3D):
        o = orc.Oracle()
        out = o.forward(bg=sc.bg.numpy(), W=cam.W, H=cam.H, means3D=inp["means3D"], dir3D=dir3D, opacities=inp["opacities"],
                        shs=shs, scales=inp["scales"], rotations=inp["rotations"], viewmatrix=cam.viewmatrix.numpy(),
                        projmatrix=cam.projmatrix.numpy(), campos=cam.campos.numpy(), tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                        kernel_size=cam.kernel_size, subpixel_offset=None, min_depth=cam.min_depth, max_depth=cam.max_depth)
        return o, float((out["color"].astype(np.float64) * gc).sum() + (out["flow"].astype(np.float64) * gf).sum())
    o, _ = fwd(inp["shs"], inp["dir3D"])
    g = o.backward(gc, zero1, gf, zero1)
    vis = np.nonzero(o.state()["tiles_touched"] > 0)[0]
    eps = 1e-2
    for i in vis[:6]:
        for (k, ch) in ((0, 0), (5, 2)):
            sp, sm = inp["shs"].copy(), inp["shs"].copy()
            sp[i, k, ch] += eps
            sm[i, k, ch] -= eps
            fd = (fwd(sp, inp["dir3D"])[1] - fwd(sm, inp["dir3D"])[1]) / (2 * eps)
            assert abs(fd - g["shs"][i, k, ch]) <= 2e-3 * max(1.0, abs(fd)) + 1e-4, (i, k, ch, fd, g["shs"][i, k, ch])
        dp, dm = inp["dir3D"].copy(), inp["dir3D"].copy()
        dp[i, 1] += eps
        dm[i, 1] -= eps
        fd = (fwd(inp["shs"], dp)[1] - fwd(inp["shs"], dm)[1]) / (2 * eps)
        assert abs(fd - g["dir3D"][i, 1]) <= 2e-3 * max(1.0, abs(fd)) + 1e-4
