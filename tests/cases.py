"""Named parity cases shared by oracle/make_golden.py and the tests."""
from __future__ import annotations

import torch

from ex4dgs_b200 import synth

# small enough for committed fixtures (~0.3 MB each), dense enough to saturate pixels
_GOLD = dict(P_static=750, P_dynamic=250, W=96, H=80, sigma_px=3.0)

GOLDEN_CASES = ["gold_base", "gold_tilted_all", "gold_subpixel", "gold_colors_precomp", "gold_cov3d_precomp",
                "gold_deg1_near"]


def make_case(name: str):
    """-> (scene, run_impl kwargs)"""
    if name == "gold_base":
        return synth.make_scene(**_GOLD), {}
    if name == "gold_tilted_all":
        sc = synth.make_scene(**_GOLD, pose="tilted", dir_nonzero=True, bg=torch.tensor([0.1, 0.5, 0.9]))
        return sc, dict(grad_kind="all")
    if name == "gold_subpixel":
        sc = synth.make_scene(**_GOLD, pose="tilted", seed=synth.SEED + 3)
        g = torch.Generator().manual_seed(5)
        return sc, dict(subpixel=torch.rand(sc.cam.H, sc.cam.W, 2, generator=g) - 0.5)
    if name == "gold_colors_precomp":
        return synth.make_scene(**_GOLD, seed=synth.SEED + 4), dict(use_colors_precomp=True, grad_kind="all")
    if name == "gold_cov3d_precomp":
        return synth.make_scene(**_GOLD, pose="tilted", seed=synth.SEED + 5), dict(use_cov3D_precomp=True)
    if name == "gold_deg1_near":
        # SH degree 1 only, a near plane that culls part of the scene, ragged image size
        sc = synth.make_scene(P_static=700, P_dynamic=0, W=101, H=67, sigma_px=3.0, seed=synth.SEED + 6,
                              pose="tilted", min_depth=4.0, max_depth=40.0)
        sc.sh_degree = 1
        return sc, dict(grad_kind="all")
    raise KeyError(name)


def make_one_tile_torture(n: int = 24000, seed: int = synth.SEED + 77):
    """>= 20k small, faint splats whose centres all fall into ONE 16x16 tile of a 48x48 image: the tile's list is
    thousands of entries long and (almost) every entry contributes to some pixel without saturating it, so the
    backward walks > 300 sub-batches of 64 (mbarrier ring wraps, n_contrib start logic) on a single tile."""
    sc = synth.make_scene(P_static=n, P_dynamic=0, W=48, H=48, sigma_px=0.7, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    W, H = sc.cam.W, sc.cam.H
    z = sc.xyz[:, 2].clone()
    pix = torch.empty(n, 2).uniform_(16.0, 32.0, generator=g)
    ndc = (2.0 * pix + 1.0) / torch.tensor([W, H], dtype=torch.float32) - 1.0
    sc.xyz = torch.stack([ndc[:, 0] * sc.cam.tanfovx * z, ndc[:, 1] * sc.cam.tanfovy * z, z], dim=1).contiguous()
    sc.xyz_disp = torch.zeros_like(sc.xyz_disp)
    op = torch.empty(n, 1).uniform_(0.008, 0.03, generator=g)
    sc.opacity = torch.log(op / (1.0 - op)).contiguous()
    sc.timestamp = 0.0
    return sc
