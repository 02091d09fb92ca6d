#!/usr/bin/env python
"""Live A/B on a GPU box: ex4dgs_b200 vs the compiled, unmodified reference (oracle/_ref).

    python tests/gpu_ab.py [--configs tiny,C1,...] [--time C3] [--out gpurun_out/ab.json]

Prints, per case, the parity figures the north star names (RGB max abs, gradient max rel,
exact equality of radii / tiles_touched / point_list / ranges / n_contrib) and optional timings.
Diagnostic script (not collected by pytest); the asserting versions live in tests/test_gpu_*.py.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _util as U  # noqa: E402
from ex4dgs_b200 import synth  # noqa: E402
from oracle import getters_oracle as GO  # noqa: E402


def compare(ours, ref, name):
    rep = {"case": name}
    for k in ("color", "depth", "acc", "flow"):
        rep["abs_" + k] = float(np.max(np.abs(ours[k] - ref[k]))) if ours[k].size else 0.0
    rep["radii_mismatch"] = int(np.sum(ours["radii"] != ref["radii"]))
    rep["idx_mismatch"] = int(np.sum(ours["idxs"] != ref["idxs"]))
    if "inter" in ours and "inter" in ref:
        a, b = ours["inter"], ref["inter"]
        rep["R"] = [int(a["R"]), int(b["R"])]
        vis = ref["radii"] > 0
        rep["P_vis"] = int(vis.sum())
        rep["tiles_touched_mismatch"] = int(np.sum(a["tiles_touched"] != b["tiles_touched"]))
        rep["depth_bits_mismatch"] = int(np.sum(a["depths"][vis].view(np.uint32) != b["depths"][vis].view(np.uint32)))
        rep["means2D_bits_mismatch"] = int(np.sum(a["means2D"][vis].view(np.uint32) != b["means2D"][vis].view(np.uint32)))
        rep["conic_opacity_maxrel"] = U.rel_err(a["conic_opacity"][vis], b["conic_opacity"][vis], 1e-6)
        rep["rgb_maxabs"] = float(np.max(np.abs(a["rgb"][vis] - b["rgb"][vis]))) if vis.any() else 0.0
        if a["R"] == b["R"]:
            rep["point_list_mismatch"] = int(np.sum(a["point_list"] != b["point_list"]))
            rep["ranges_mismatch"] = int(np.sum(a["ranges"] != b["ranges"]))
            rep["n_contrib_mismatch"] = int(np.sum(a["n_contrib"] != b["n_contrib"]))
            rep["final_T_maxabs"] = float(np.max(np.abs(a["final_T"] - b["final_T"])))
        rep["color_bits_mismatch"] = int(np.sum(ours["color"].view(np.uint32) != ref["color"].view(np.uint32)))
    if "grads" in ours and "grads" in ref:
        for k in ref["grads"]:
            fl = U.grad_floor(ref["grads"][k])
            rep["grad_rel_" + k] = U.rel_err(ours["grads"][k], ref["grads"][k], fl)
            rep["grad_floor_" + k] = fl
    return rep


def time_impl(mod, sc, iters=10, warmup=3, backward=True, dev="cuda"):
    inp = {k: v.to(dev).requires_grad_(backward) for k, v in GO.flat_inputs(sc).items()}
    P = inp["means3D"].shape[0]
    means2D = torch.zeros(P, 3, device=dev, requires_grad=backward)
    rs = U.settings_for(mod, sc, dev)
    go = {k: v.to(dev) for k, v in synth.grad_outputs(sc).items()}
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    ts_f, ts_b = [], []
    R = -1
    for it in range(warmup + iters):
        flush.fill_(1.0)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        rast = mod.GaussianRasterizer(rs)
        e0.record()
        color, radii, depth, flow, acc, idxs = rast(means3D=inp["means3D"], means2D=means2D, dir3D=inp["dir3D"],
                                                    opacities=inp["opacities"], shs=inp["shs"], scales=inp["scales"],
                                                    rotations=inp["rotations"])
        e1.record()
        if backward:
            torch.autograd.backward([color, depth, flow, acc], [go["grad_color"], go["grad_depth"], go["grad_flow"], go["grad_acc"]])
            for v in list(inp.values()) + [means2D]:
                v.grad = None
        e2.record()
        torch.cuda.synchronize()
        if it >= warmup:
            ts_f.append(e0.elapsed_time(e1))
            ts_b.append(e1.elapsed_time(e2))
        if color.grad_fn is not None:
            R = int(color.grad_fn.num_rendered)
    return dict(fwd_ms=float(np.median(ts_f)), bwd_ms=float(np.median(ts_b)), R=R, P=P,
                P_vis=int((radii > 0).sum().item()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="tiny,C1,C1d")
    ap.add_argument("--time", default="")
    ap.add_argument("--out", default="gpurun_out/ab.json")
    ap.add_argument("--cull", type=int, default=0)
    args = ap.parse_args()
    ref = U.reference_module()
    ours = U.ours_module()
    ours.set_default_flags(bool(args.cull))
    print("reference module:", "available" if ref else "MISSING", flush=True)
    reports = []
    variants = [
        ("base", dict(), dict()),
        ("tilted+dir+bg", dict(pose="tilted", dir_nonzero=True, bg=torch.tensor([0.1, 0.5, 0.9])), dict(grad_kind="all")),
        ("subpixel", dict(pose="tilted"), dict(subpixel=True)),
        ("colors_precomp", dict(), dict(use_colors_precomp=True)),
        ("cov3D_precomp", dict(pose="tilted"), dict(use_cov3D_precomp=True)),
    ]
    for cfg in [c for c in args.configs.split(",") if c]:
        for vname, skw, rkw in variants:
            if cfg not in ("tiny", "C1", "C1d") and vname != "base":
                continue
            sc = synth.make_config(cfg, **skw)
            rkw = dict(rkw)
            if rkw.pop("subpixel", False):
                g = torch.Generator().manual_seed(5)
                rkw["subpixel"] = torch.rand(sc.cam.H, sc.cam.W, 2, generator=g) - 0.5
            o = U.run_impl(ours, sc, is_ref=False, **rkw)
            name = "%s/%s" % (cfg, vname)
            if ref is not None:
                r = U.run_impl(ref, sc, is_ref=True, **rkw)
                rep = compare(o, r, name)
            else:
                rep = {"case": name, "note": "no reference", "R": int(o["inter"]["R"])}
            print(json.dumps(rep), flush=True)
            reports.append(rep)
    for cfg in [c for c in args.time.split(",") if c]:
        sc = synth.make_config(cfg)
        t_o = time_impl(ours, sc)
        rep = {"case": "time/" + cfg, "ours": t_o}
        if ref is not None:
            rep["ref"] = time_impl(ref, sc)
        print(json.dumps(rep), flush=True)
        reports.append(rep)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(reports, f, indent=1)


if __name__ == "__main__":
    main()
