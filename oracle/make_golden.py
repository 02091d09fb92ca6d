#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference (oracle/_ref).

Must run on a GPU box (the reference is CUDA-only, rasterize_points.cu:80):

    gpurun -- python oracle/make_golden.py --out gpurun_out/golden
    cp gpurun_out/golden/*.npz tests/golden/

Every case is a seeded scene of ex4dgs_b200.synth (inputs are regenerated from the seed by the
tests, only the reference's outputs are stored): the six outputs, the tile-list intermediates
(depths, means2D, conic_opacity, rgb, clamped, tiles_touched, 64-bit sorted keys, point_list, ranges,
final_T, n_contrib) and all gradients.  These files pin the CPU oracle (tests/test_oracle_golden.py)
and are a second, GPU-box-independent reference for the CUDA path (tests/test_gpu_parity.py).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import _util as U  # noqa: E402
from tests.cases import GOLDEN_CASES, make_case  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/golden")
    args = ap.parse_args()
    ref = U.reference_module()
    if ref is None or not torch.cuda.is_available():
        raise SystemExit("needs oracle/_ref (python oracle/build_ref.py) and a GPU")
    os.makedirs(args.out, exist_ok=True)
    for name in GOLDEN_CASES:
        sc, rkw = make_case(name)
        r = U.run_impl(ref, sc, kind="ref", **rkw)
        flat = {k: r[k] for k in ("color", "radii", "depth", "flow", "acc", "idxs")}
        for k, v in r["inter"].items():
            if k in ("cov3D",):
                continue
            flat["inter_" + k] = np.asarray(v)
        for k, v in r["grads"].items():
            flat["grad_" + k] = v
        path = os.path.join(args.out, name + ".npz")
        np.savez_compressed(path, **flat)
        print(name, "R=%d" % r["inter"]["R"], "P_vis=%d" % int((r["radii"] > 0).sum()), "->", path,
              "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
