"""Generates tests/golden/ply_fixture/{point_cloud.ply,dynamic_point_cloud.ply,tensors.npz} by running the
reference's OWN CGaussianModel.save_ply / load_ply (/root/reference/scene/c_gaussian_model.py:514-672,
imported unmodified) on a small seeded model.  Test infrastructure; run in the build container only
(the GPU box has no /root/reference):

    python oracle/make_ply_golden.py

Two of the reference's imports do not exist in this image and are replaced by minimal stand-ins
before the import: `simple_knn._C` (unused by save/load) and `plyfile` - a ~40-line container that
writes/reads exactly the binary_little_endian layout plyfile produces for a structured array of 'f4'
fields (header `property float <name>` per field, packed little-endian rows).  The reference code
decides WHAT goes into which named column and how columns map back to tensors; that mapping is what
ex4dgs_b200/model_io.py must reproduce.  load_ply hard-codes device="cuda"; torch.tensor is wrapped
to drop the device argument for the duration of the call.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "ply_fixture")


# ---- stand-in for the `plyfile` package -------------------------------------------------------
class _Prop:
    def __init__(self, name):
        self.name = name


class PlyElement:
    def __init__(self, data, name):
        self.data, self.name = data, name
        self.properties = [_Prop(n) for n in data.dtype.names]

    @staticmethod
    def describe(data, name):
        assert all(data.dtype[n] == np.dtype("f4") for n in data.dtype.names)
        return PlyElement(data, name)

    def __getitem__(self, key):
        return self.data[key]


class PlyData:
    def __init__(self, elements):
        self.elements = elements

    def write(self, path):
        el = self.elements[0]
        hdr = ["ply", "format binary_little_endian 1.0", "element %s %d" % (el.name, len(el.data))]
        hdr += ["property float %s" % n for n in el.data.dtype.names] + ["end_header"]
        with open(path, "wb") as f:
            f.write(("\n".join(hdr) + "\n").encode("ascii"))
            f.write(el.data.astype(el.data.dtype.newbyteorder("<")).tobytes())

    @staticmethod
    def read(path):
        with open(path, "rb") as f:
            assert f.readline() == b"ply\n"
            names, n = [], None
            while True:
                t = f.readline().decode().split()
                if t[0] == "element":
                    n = int(t[2])
                elif t[0] == "property":
                    assert t[1] == "float"
                    names.append(t[2])
                elif t[0] == "end_header":
                    break
            data = np.frombuffer(f.read(), dtype=np.dtype([(x, "<f4") for x in names]), count=n)
        return PlyData([PlyElement(data, "vertex")])


def _install_stubs():
    m = types.ModuleType("plyfile")
    m.PlyData, m.PlyElement = PlyData, PlyElement
    sys.modules["plyfile"] = m
    pkg = types.ModuleType("simple_knn")
    sub = types.ModuleType("simple_knn._C")
    sub.distCUDA2 = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("not available"))
    pkg._C = sub
    sys.modules["simple_knn"] = pkg
    sys.modules["simple_knn._C"] = sub


TENSORS = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "_xyz_disp", "_xyz_motion",
           "_features_dc_motion", "_features_rest_motion", "_scaling_motion", "_opacity_motion", "_opacity_duration_center",
           "_opacity_duration_var", "_rotation_motion")


def main():
    _install_stubs()
    sys.path.insert(0, "/root/reference")
    # the file itself, unmodified; loaded by path because `import scene` pulls in the dataset readers
    # (natsort, PIL, ... absent here)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_c_gaussian_model", "/root/reference/scene/c_gaussian_model.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    CGaussianModel = mod.CGaussianModel

    g = torch.Generator().manual_seed(20240925)
    Ns, Nd, sh_degree = 23, 11, 3
    duration, interval, time_pad = 40, 10, 2
    time_shift = time_pad + interval                     # interp_type "cube" (c_gaussian_model.py:76,119)
    K = int(np.ceil((duration + time_shift + time_pad * 2 + 1) / interval)) + 1 + 4

    def R(*s):
        return torch.randn(*s, generator=g)

    src = {"_xyz": R(Ns, 3), "_features_dc": R(Ns, 1, 3), "_features_rest": R(Ns, 15, 3), "_opacity": R(Ns, 1),
           "_scaling": R(Ns, 3), "_rotation": R(Ns, 4), "_xyz_disp": R(Ns, 3), "_xyz_motion": R(Nd, K, 3),
           "_features_dc_motion": R(Nd, 1, 3), "_features_rest_motion": R(Nd, 15, 3), "_scaling_motion": R(Nd, 3),
           "_opacity_motion": R(Nd, 1), "_opacity_duration_center": R(Nd, 2, 1), "_opacity_duration_var": R(Nd, 2, 1),
           "_rotation_motion": R(Nd, K, 4)}

    m = CGaussianModel.__new__(CGaussianModel)           # __init__ calls .cuda(); save/load need none of it
    for k, v in src.items():
        setattr(m, k, v.clone())
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "point_cloud.ply")
    m.save_ply(path)                                      # the reference writes both files

    # the reference reads them back
    m2 = CGaussianModel.__new__(CGaussianModel)
    m2.max_sh_degree, m2.duration, m2.interval = sh_degree, duration, interval
    m2.time_shift, m2.time_pad, m2.motion_degree, m2.opacity_degree = time_shift, time_pad, 1, 2
    real_tensor = torch.tensor
    torch.tensor = lambda data, **kw: real_tensor(data, **{k: v for k, v in kw.items() if k != "device"})
    try:
        m2.load_ply(path)
    finally:
        torch.tensor = real_tensor
    assert m2.keyframe_num == K
    out = {}
    for k in TENSORS:
        a, b = src[k], getattr(m2, k).detach()
        assert a.shape == b.shape and torch.equal(a, b), k  # the reference round-trips its own format exactly
        out[k] = b.numpy()
    out["meta"] = np.array([sh_degree, duration, interval, time_pad, time_shift, K], dtype=np.float64)
    out["static_names"] = np.array(m.construct_list_of_static_attributes())
    out["dynamic_names"] = np.array(m.construct_list_of_dynamic_attributes())
    np.savez_compressed(os.path.join(OUT, "tensors.npz"), **out)
    print("wrote", OUT, {k: os.path.getsize(os.path.join(OUT, k)) for k in os.listdir(OUT)})


if __name__ == "__main__":
    main()
