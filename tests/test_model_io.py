"""CPU tests of ex4dgs_b200/model_io.py against the fixture written and read back by the
reference's own CGaussianModel.save_ply / load_ply (oracle/make_ply_golden.py -> tests/golden/ply_fixture)."""
import os

import numpy as np
import pytest
import torch

from ex4dgs_b200 import model_io as mio

FIX = os.path.join(os.path.dirname(__file__), "golden", "ply_fixture")


def _fixture():
    z = np.load(os.path.join(FIX, "tensors.npz"))
    sh_degree, duration, interval, time_pad, time_shift, K = [float(x) for x in z["meta"]]
    return z, dict(sh_degree=int(sh_degree), duration=duration, interval=interval, time_pad=time_pad), time_shift, int(K)


def test_load_matches_reference_load_ply():
    z, kw, time_shift, K = _fixture()
    m = mio.load_model(os.path.join(FIX, "point_cloud.ply"), **kw)
    for name in mio.STATIC_TENSORS + mio.DYNAMIC_TENSORS:
        got = getattr(m, name)
        assert got.dtype == torch.float32 and got.is_contiguous()
        assert tuple(got.shape) == z[name].shape, name
        assert np.array_equal(got.numpy(), z[name]), name           # bit-exact: pure data movement
    assert m.keyframe_num == K and m.time_shift == time_shift
    assert m.get_features().shape == (m.num_static + m.num_dynamic, 16, 3)
    assert torch.equal(m.get_features()[: m.num_static, :1], m._features_dc)


def test_save_is_byte_identical_to_reference_save_ply(tmp_path):
    z, kw, _, _ = _fixture()
    m = mio.load_model(os.path.join(FIX, "point_cloud.ply"), **kw)
    out = str(tmp_path / "iteration_7" / "point_cloud.ply")
    mio.save_model(m, out)
    for f in ("point_cloud.ply", "dynamic_point_cloud.ply"):
        a = open(os.path.join(FIX, f), "rb").read()
        b = open(os.path.join(os.path.dirname(out), f), "rb").read()
        assert a == b, f


def test_attribute_lists_match_reference():
    z, _, _, K = _fixture()
    assert mio.static_attributes(45) == [str(x) for x in z["static_names"]]
    assert mio.dynamic_attributes(K, 45, 3, 2) == [str(x) for x in z["dynamic_names"]]


def test_ply_container_variants(tmp_path):
    names = ["x", "y", "scale_0", "scale_1"]
    data = np.arange(12, dtype=np.float32).reshape(3, 4) * 0.5
    p = str(tmp_path / "a.ply")
    mio.write_ply(p, names, data)
    n2, d2 = mio.read_ply(p)
    assert n2 == names and np.array_equal(d2, data)
    # ascii and big-endian / mixed-type bodies (other writers), comments in the header
    q = str(tmp_path / "b.ply")
    with open(q, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment hi\nelement vertex 3\n" + "".join("property float %s\n" % n for n in names) + "end_header\n")
        for r in data:
            f.write(" ".join(repr(float(v)) for v in r) + "\n")
    n3, d3 = mio.read_ply(q)
    assert n3 == names and np.array_equal(d3, data)
    r = str(tmp_path / "c.ply")
    with open(r, "wb") as f:
        f.write(b"ply\nformat binary_big_endian 1.0\nelement vertex 3\nproperty double x\nproperty uchar y\nend_header\n")
        rec = np.zeros(3, dtype=np.dtype([("x", ">f8"), ("y", "u1")]))
        rec["x"] = [1.5, 2.5, 3.5]
        rec["y"] = [7, 8, 9]
        f.write(rec.tobytes())
    n4, d4 = mio.read_ply(r)
    assert n4 == ["x", "y"] and np.array_equal(d4, np.array([[1.5, 7], [2.5, 8], [3.5, 9]], dtype=np.float32))
    # empty model
    e = str(tmp_path / "e.ply")
    mio.write_ply(e, names, np.zeros((0, 4), np.float32))
    n5, d5 = mio.read_ply(e)
    assert n5 == names and d5.shape == (0, 4)


def test_errors(tmp_path):
    z, kw, _, _ = _fixture()
    with pytest.raises(ValueError):
        mio.read_ply(os.path.join(FIX, "tensors.npz"))               # not a PLY
    bad = dict(kw)
    bad["duration"] = kw["duration"] + 50                             # implies a different keyframe count
    with pytest.raises(ValueError):
        mio.load_model(os.path.join(FIX, "point_cloud.ply"), **bad)
    with pytest.raises(ValueError):
        mio.load_model(os.path.join(FIX, "point_cloud.ply"), **dict(kw, sh_degree=2))   # f_rest count mismatch
    # truncated body
    src = open(os.path.join(FIX, "point_cloud.ply"), "rb").read()
    t = str(tmp_path / "point_cloud.ply")
    open(t, "wb").write(src[:-17])
    with pytest.raises(ValueError):
        mio.read_ply(t)
    with pytest.raises(ValueError):
        mio.save_model(mio.load_model(os.path.join(FIX, "point_cloud.ply"), **kw), str(tmp_path / "model.ply"))


def test_capture_round_trip(tmp_path):
    z, kw, time_shift, K = _fixture()
    m = mio.load_model(os.path.join(FIX, "point_cloud.ply"), **kw)
    # a checkpoint as train.py:197 writes it: (capture(), iteration); statistics slots are arbitrary tensors here
    cap = list(mio.to_capture(m))
    slots = dict(zip(mio.CAPTURE_SLOTS, range(len(mio.CAPTURE_SLOTS))))
    cap[slots["max_radii2D"]] = torch.arange(m.num_static, dtype=torch.float32)
    cap[slots["optimizer_state"]] = {"state": {}, "param_groups": [{"name": "xyz", "lr": 1e-4}]}
    cap[slots["spatial_lr_scale"]] = 2.5
    p = str(tmp_path / "chkpnt30000.pth")
    torch.save((tuple(cap), 30000), p)
    m2, it = mio.load_checkpoint(p, time_pad=kw["time_pad"])
    assert it == 30000 and m2.keyframe_num == K and m2.time_shift == time_shift
    for name in mio.STATIC_TENSORS + mio.DYNAMIC_TENSORS:
        assert torch.equal(getattr(m2, name), getattr(m, name)), name
    assert m2.extras["spatial_lr_scale"] == 2.5 and torch.equal(m2.extras["max_radii2D"], cap[slots["max_radii2D"]])
    cap2 = mio.to_capture(m2)
    assert len(cap2) == len(mio.CAPTURE_SLOTS) and cap2[slots["optimizer_state"]]["param_groups"][0]["name"] == "xyz"
    with pytest.raises(ValueError):
        mio.from_capture(tuple(cap[:-1]))


def test_load_iteration_lookup(tmp_path):
    z, kw, _, _ = _fixture()
    m = mio.load_model(os.path.join(FIX, "point_cloud.ply"), **kw)
    for it in (7000, 30000):
        mio.save_model(m, str(tmp_path / "point_cloud" / ("iteration_%d" % it) / "point_cloud.ply"))
    m2, it = mio.load_iteration(str(tmp_path), -1, **kw)
    assert it == 30000 and torch.equal(m2._xyz_motion, m._xyz_motion)
    m3, it3 = mio.load_iteration(str(tmp_path), 7000, **kw)
    assert it3 == 7000 and torch.equal(m3._rotation, m._rotation)
