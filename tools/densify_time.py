#!/usr/bin/env python
"""Time CGaussianModel.prune_points and densification_postfix at workload C3 (1.5 M static + 0.5 M dynamic Gaussians, 36
keyframes: 247 M parameters + two moments each + 18 statistics tensors) - the reference's own methods against
ex4dgs_b200.densify bound onto the same class - and print one JSON line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ex4dgs_b200 import densify, synth  # noqa: E402
from tests.test_gpu_densify import _model  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


def main():
    cls = bench.load_reference_model_class()
    sc = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "C3")
    out = {}
    for arm in ("reference", "ours"):
        res = {"prune_points_ms": [], "densification_postfix_ms": []}
        for rep in range(3):
            m = _model(sc, cls, steps=1)
            if arm == "ours":
                densify.install(m)
            g = torch.Generator(device="cuda").manual_seed(rep)
            ms = torch.rand(m._xyz.shape[0], generator=g, device="cuda") < 0.1
            md = torch.rand(m._xyz_motion.shape[0], generator=g, device="cuda") < 0.1
            res["prune_points_ms"].append(timed(lambda: m.prune_points(ms, md)))
            sel_s = torch.nonzero(torch.rand(m._xyz.shape[0], generator=g, device="cuda") < 0.05).squeeze(1)
            sel_d = torch.nonzero(torch.rand(m._xyz_motion.shape[0], generator=g, device="cuda") < 0.05).squeeze(1)
            names_s = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "_xyz_disp")
            names_d = ("_xyz_motion", "_features_dc_motion", "_features_rest_motion", "_scaling_motion", "_opacity_motion",
                       "_opacity_duration_center", "_opacity_duration_var", "_rotation_motion")
            new = [getattr(m, n)[sel_s] for n in names_s] + [getattr(m, n)[sel_d] for n in names_d]
            res["densification_postfix_ms"].append(timed(lambda: m.densification_postfix(*new)))
            del m, new
            torch.cuda.empty_cache()
        out[arm] = {k: min(v) for k, v in res.items()}
    out["speedup"] = {k: out["reference"][k] / out["ours"][k] for k in out["ours"]}
    out["what"] = "host wall-clock incl. synchronisation, best of 3; 10 % of the Gaussians pruned, then 5 % appended"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
