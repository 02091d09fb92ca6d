"""The model-side getters restated in ex4dgs_b200/synth.py against the reference's own Python
(utils/interpolations.py, imported from /root/reference when present - it cannot travel to the GPU
box) and against a committed fixture generated from it (tests/golden/interp_fixture.npz,
generator: this file's __main__)."""
import os
import sys

import numpy as np
import pytest
import torch

from ex4dgs_b200 import synth
from oracle import getters_oracle as GO  # noqa: E402

FIX = os.path.join(os.path.dirname(__file__), "golden", "interp_fixture.npz")
REF = "/root/reference"


def _scene():
    return synth.make_scene(40, 60, 64, 48, seed=21)


def _ours(sc, t):
    k, d = GO.frame_indices(sc, t)
    tau = (t + sc.time_shift) / sc.interval
    return dict(xyz=GO.cube_interp(sc.xyz_motion, k, d).numpy(),
                rot=GO.quat_slerp(sc.rotation_motion[:, k], sc.rotation_motion[:, k + 1], d).numpy(),
                opa=GO.time_bigaussian(sc.opacity_center, sc.opacity_var, tau, sc.var_pad / sc.interval).numpy())


def _reference(sc, t):
    sys.path.insert(0, REF)
    from utils.interpolations import cube_interpolate, quat_slerp_interp_uniiterval, time_bigaussian
    k, d = GO.frame_indices(sc, t)
    y = sc.xyz_motion
    tau = (t + sc.time_shift) / sc.interval
    return dict(xyz=cube_interpolate(y[:, k - 1, :3], y[:, k, :3], y[:, k + 1, :3], y[:, k + 2, :3], d).numpy(),
                rot=quat_slerp_interp_uniiterval(sc.rotation_motion[:, k], sc.rotation_motion[:, k + 1], d).numpy(),
                opa=time_bigaussian(sc.opacity_center, sc.opacity_var, tau, var_min=sc.var_pad / sc.interval).numpy())


TS = [0.0, 7.5, 137.0, 299.0]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_getters_match_reference_python():
    sc = _scene()
    for t in TS:
        a, b = _ours(sc, t), _reference(sc, t)
        for k in a:
            assert np.array_equal(a[k], b[k]), (t, k)


def test_getters_match_committed_fixture():
    if not os.path.exists(FIX):
        pytest.skip("fixture not generated")
    f = np.load(FIX)
    sc = _scene()
    for i, t in enumerate(TS):
        a = _ours(sc, t)
        for k in a:
            assert np.allclose(a[k], f["%s_%d" % (k, i)], rtol=1e-6, atol=1e-7), (t, k)


def test_flat_inputs_layout_static_first():
    sc = _scene()
    inp = GO.flat_inputs(sc)
    ns = sc.xyz.shape[0]
    assert inp["means3D"].shape == (sc.P, 3) and inp["shs"].shape == (sc.P, 16, 3)
    assert torch.equal(inp["rotations"][:ns], sc.rotation)          # static quats are passed raw (c_gaussian_model.py:198)
    assert torch.allclose(inp["rotations"][ns:].norm(dim=1), torch.ones(sc.P - ns), atol=1e-5)
    assert torch.equal(inp["shs"][ns:], sc.features_motion)


if __name__ == "__main__":           # regenerate the fixture from the reference's Python
    sc = _scene()
    out = {}
    for i, t in enumerate(TS):
        for k, v in _reference(sc, t).items():
            out["%s_%d" % (k, i)] = v
    np.savez_compressed(FIX, **out)
    print("wrote", FIX)
