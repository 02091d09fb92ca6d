#!/usr/bin/env python
"""Run-to-run spread of the small-scene gradient comparison against the (deterministic) CPU oracle:
the scene of tests/test_gpu_parity.py::test_nine_coefficient_sh_rows_against_oracle, N runs per cull mode."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ex4dgs_b200 import synth  # noqa: E402
from oracle import getters_oracle as GO  # noqa: E402
from tests import _util as U  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    mod, orc = U.ours_module(), U.oracle_module()
    sc = synth.make_scene(900, 300, 112, 80, sigma_px=3.0, seed=synth.SEED + 41, pose="tilted")
    sc.sh_degree = 2
    go = synth.grad_outputs(sc)

    def run(m, dev):
        inp = {k: v.to(dev) for k, v in GO.flat_inputs(sc).items()}
        inp["shs"] = inp["shs"][:, :9].contiguous()
        t = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
        m2 = torch.zeros(t["means3D"].shape[0], 3, device=dev, requires_grad=True)
        out = m.GaussianRasterizer(U.settings_for(m, sc, dev))(means3D=t["means3D"], means2D=m2, dir3D=t["dir3D"], opacities=t["opacities"],
                                                             shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
        torch.autograd.backward([out[0], out[2], out[3], out[4]], [go[k].to(dev) for k in ("grad_color", "grad_depth", "grad_flow", "grad_acc")])
        res = {k: v.grad.detach().cpu().numpy() for k, v in t.items()}
        res["means2D"] = m2.grad.detach().cpu().numpy()
        return res

    gb = run(orc, "cpu")
    for cull in (0, 1):
        mod.set_default_flags(bool(cull))
        errs = {k: [] for k in gb}
        for _ in range(n):
            ga = run(mod, "cuda")
            for k, g in gb.items():
                errs[k].append(U.rel_err(ga[k], g, U.grad_floor(g)))
        for k, v in errs.items():
            v = np.array(v)
            print("cull=%d %-10s min %.2e median %.2e max %.2e  runs>1e-3: %d/%d" % (cull, k, v.min(), np.median(v), v.max(), int((v > 1e-3).sum()), n))


if __name__ == "__main__":
    main()
