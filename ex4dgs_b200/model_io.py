"""Trained-model I/O: the reference's two-file PLY format and its checkpoint
tuple, loaded straight into the tensors the rasterizer front-end consumes.

Mirrors, name for name and layout for layout:

  save_ply   scene/c_gaussian_model.py:514-546   point_cloud.ply (static) + dynamic_point_cloud.ply
  load_ply   scene/c_gaussian_model.py:558-672
  capture / restore (the tensor slots)  scene/c_gaussian_model.py:217-320, written by train.py:197
  attribute lists  construct_list_of_static_attributes / _dynamic_attributes  :473-512

The reference goes through the `plyfile` package (absent from this image) and fills every column
with a Python loop over property names; here the binary PLY body is read as ONE numpy structured
array (a single contiguous read of N x n_props float32) and split with reshapes/transposes, then
moved to the device once per tensor.  Pure host code: no CUDA needed, covered by the CPU tests
(tests/test_model_io.py) against a fixture produced by the reference's own save_ply/load_ply code
(oracle/make_ply_golden.py).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1",
    "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
    "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
    "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


# ---------------------------------------------------------------------------------------------
# PLY container (single `vertex` element of scalar properties - all the reference ever writes)
# ---------------------------------------------------------------------------------------------
def read_ply(path: str) -> Tuple[List[str], np.ndarray]:
    """-> (property names, float32 array [N, n_props]).  binary_little_endian (what plyfile writes
    by default and the reference uses), binary_big_endian and ascii are accepted."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError("%s: not a PLY file" % path)
        fmt = None
        count = None
        names: List[str] = []
        types: List[str] = []
        in_vertex = False
        while True:
            line = f.readline()
            if not line:
                raise ValueError("%s: unterminated PLY header" % path)
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                if count is not None and not in_vertex:
                    pass
                in_vertex = (tok[1] == "vertex") and count is None
                if in_vertex:
                    count = int(tok[2])
                elif int(tok[2]) != 0 and count is None:
                    raise ValueError("%s: element '%s' before 'vertex' is not supported" % (path, tok[1]))
            elif tok[0] == "property":
                if not in_vertex:
                    continue
                if tok[1] == "list":
                    raise ValueError("%s: list properties are not supported" % path)
                if tok[1] not in _PLY_TYPES:
                    raise ValueError("%s: unknown property type %s" % (path, tok[1]))
                types.append(_PLY_TYPES[tok[1]])
                names.append(tok[2])
            elif tok[0] == "end_header":
                break
        if fmt is None or count is None:
            raise ValueError("%s: header lacks format/vertex element" % path)
        if fmt == "ascii":
            data = np.loadtxt(f, dtype=np.float64, max_rows=count, ndmin=2) if count else np.zeros((0, len(names)))
            if data.shape != (count, len(names)):
                raise ValueError("%s: expected %d x %d ascii values, got %s" % (path, count, len(names), data.shape))
            return names, data.astype(np.float32)
        if fmt not in ("binary_little_endian", "binary_big_endian"):
            raise ValueError("%s: unknown PLY format %s" % (path, fmt))
        order = "<" if fmt == "binary_little_endian" else ">"
        if all(t == "f4" for t in types):
            # the reference's files: one flat read, no per-column work
            raw = np.fromfile(f, dtype=order + "f4", count=count * len(names))
            if raw.size != count * len(names):
                raise ValueError("%s: truncated body (%d of %d values)" % (path, raw.size, count * len(names)))
            return names, raw.reshape(count, len(names)).astype(np.float32, copy=False)
        dt = np.dtype([(n, order + t) for n, t in zip(names, types)])
        rec = np.fromfile(f, dtype=dt, count=count)
        if rec.shape[0] != count:
            raise ValueError("%s: truncated body" % path)
        out = np.empty((count, len(names)), dtype=np.float32)
        for i, n in enumerate(names):
            out[:, i] = rec[n]
        return names, out


def write_ply(path: str, names: Sequence[str], data: np.ndarray) -> None:
    """Write [N, n_props] float32 as a binary_little_endian PLY with one `vertex` element of `float`
    properties - byte for byte what PlyData([PlyElement.describe(arr_f4, 'vertex')]).write() emits."""
    data = np.ascontiguousarray(data, dtype="<f4")
    if data.ndim != 2 or data.shape[1] != len(names):
        raise ValueError("data must be [N, %d]" % len(names))
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    header = ["ply", "format binary_little_endian 1.0", "element vertex %d" % data.shape[0]]
    header += ["property float %s" % n for n in names]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        data.tofile(f)


# ---------------------------------------------------------------------------------------------
# attribute lists (c_gaussian_model.py:473-512)
# ---------------------------------------------------------------------------------------------
def static_attributes(n_rest: int = 45) -> List[str]:
    l = ["x", "y", "z", "nx", "ny", "nz"]
    l += ["f_dc_%d" % i for i in range(3)]
    l += ["f_rest_%d" % i for i in range(n_rest)]
    l.append("opacity")
    l += ["scale_%d" % i for i in range(3)]
    l += ["rot_%d" % i for i in range(4)]
    l += ["xyz_disp_%d" % i for i in range(3)]
    return l


def dynamic_attributes(K: int, n_rest: int = 45, motion_width: int = 3, opacity_degree: int = 2) -> List[str]:
    l = ["motion_xyz_%d_%d" % (i, j) for i in range(K) for j in range(motion_width)]
    l += ["motion_f_dc_%d" % i for i in range(3)]
    l += ["motion_f_rest_%d" % i for i in range(n_rest)]
    l += ["motion_scale_%d" % i for i in range(3)]
    l.append("motion_opacity")
    l += ["motion_opacity_c_%d" % i for i in range(opacity_degree)]
    l += ["motion_opacity_v_%d" % i for i in range(opacity_degree)]
    l += ["motion_rot_%d_%d" % (i, j) for i in range(K) for j in range(4)]
    return l


# ---------------------------------------------------------------------------------------------
# the model's tensors under the reference's attribute names
# ---------------------------------------------------------------------------------------------
STATIC_TENSORS = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "_xyz_disp")
DYNAMIC_TENSORS = ("_xyz_motion", "_features_dc_motion", "_features_rest_motion", "_scaling_motion", "_opacity_motion",
                   "_opacity_duration_center", "_opacity_duration_var", "_rotation_motion")


@dataclass
class GaussianArrays:
    """The subset of CGaussianModel state the render path reads (same attribute names, shapes and
    activations: raw log-scales, raw logit opacities, un-normalised static quaternions), usable with
    ex4dgs_b200.frontend.FusedGetters and the reference's unmodified render()."""
    _xyz: torch.Tensor                     # [Ns,3]
    _features_dc: torch.Tensor             # [Ns,1,3]
    _features_rest: torch.Tensor           # [Ns,15,3]
    _opacity: torch.Tensor                 # [Ns,1]
    _scaling: torch.Tensor                 # [Ns,3]
    _rotation: torch.Tensor                # [Ns,4]
    _xyz_disp: torch.Tensor                # [Ns,3]
    _xyz_motion: torch.Tensor              # [Nd,K,3]
    _features_dc_motion: torch.Tensor      # [Nd,1,3]
    _features_rest_motion: torch.Tensor    # [Nd,15,3]
    _scaling_motion: torch.Tensor          # [Nd,3]
    _opacity_motion: torch.Tensor          # [Nd,1]
    _opacity_duration_center: torch.Tensor  # [Nd,2,1]
    _opacity_duration_var: torch.Tensor    # [Nd,2,1]
    _rotation_motion: torch.Tensor         # [Nd,K,4]
    max_sh_degree: int = 3
    active_sh_degree: int = 3
    duration: float = 1.0
    interval: float = 1.0
    time_pad: float = 1.0
    time_shift: float = 1.0                # time_pad (+ interval for "cube"/"pchip": c_gaussian_model.py:76,119)
    var_pad: float = 3.0
    kernel_size: float = 0.1
    interp_type: str = "cube"
    rot_interp_type: str = "slerp"
    extras: Dict[str, object] = field(default_factory=dict)   # checkpoint-only slots (statistics, optimizer state)

    @property
    def keyframe_num(self) -> int:
        return int(self._xyz_motion.shape[1]) if self._xyz_motion.dim() == 3 else 0

    @property
    def num_static(self) -> int:
        return int(self._xyz.shape[0])

    @property
    def num_dynamic(self) -> int:
        return int(self._xyz_motion.shape[0])

    def get_features(self, mode: int = 0) -> torch.Tensor:
        """c_gaussian_model.py:337-353."""
        s = torch.cat((self._features_dc, self._features_rest), dim=1)
        d = torch.cat((self._features_dc_motion, self._features_rest_motion), dim=1)
        if mode == 1:
            return s
        if mode == 2:
            return d
        return torch.cat((s, d), dim=0)

    def to(self, device) -> "GaussianArrays":
        kw = {n: getattr(self, n).to(device) for n in STATIC_TENSORS + DYNAMIC_TENSORS}
        rest = {k: getattr(self, k) for k in ("max_sh_degree", "active_sh_degree", "duration", "interval", "time_pad",
                                              "time_shift", "var_pad", "kernel_size", "interp_type", "rot_interp_type", "extras")}
        return GaussianArrays(**kw, **rest)


def expected_keyframes(duration: float, interval: float, time_pad: float, time_shift: float) -> int:
    """keyframe_num as load_ply recomputes it (c_gaussian_model.py:601)."""
    return math.ceil((duration + time_shift + time_pad * 2 + 1) / interval) + 1 + 4


def _time_shift(time_pad: float, interval: float, interp_type: str) -> float:
    return time_pad + (interval if interp_type in ("cube", "pchip") else 0.0)


# ---------------------------------------------------------------------------------------------
# save / load (PLY pair)
# ---------------------------------------------------------------------------------------------
def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().to("cpu", torch.float32).numpy()


def save_model(m: GaussianArrays, path: str) -> None:
    """Write `path` (…/point_cloud.ply) and its sibling dynamic_point_cloud.ply exactly like
    CGaussianModel.save_ply (c_gaussian_model.py:514-546): SH blocks are stored channel-major
    (transpose(1,2).flatten), `nx ny nz` are zeros, keyframes are flattened frame-major."""
    if not path.endswith("point_cloud.ply"):
        raise ValueError("path must end with point_cloud.ply (the dynamic file name is derived from it)")
    Ns, Nd = m.num_static, m.num_dynamic
    xyz = _np(m._xyz).reshape(Ns, 3)
    cols = [xyz, np.zeros_like(xyz),
            _np(m._features_dc.transpose(1, 2).flatten(start_dim=1)) if Ns else np.zeros((0, 3), np.float32),
            _np(m._features_rest.transpose(1, 2).flatten(start_dim=1)) if Ns else np.zeros((0, 3 * m._features_rest.shape[1]), np.float32),
            _np(m._opacity).reshape(Ns, 1), _np(m._scaling).reshape(Ns, 3), _np(m._rotation).reshape(Ns, 4),
            _np(m._xyz_disp).reshape(Ns, 3)]
    n_rest = int(m._features_rest.shape[1] * m._features_rest.shape[2])
    write_ply(path, static_attributes(n_rest), np.concatenate(cols, axis=1))

    K = m.keyframe_num
    n_rest_m = int(m._features_rest_motion.shape[1] * m._features_rest_motion.shape[2])
    width = int(m._xyz_motion.shape[2]) if m._xyz_motion.dim() == 3 else 3
    deg = int(m._opacity_duration_center.shape[1]) if m._opacity_duration_center.dim() >= 2 else 2
    cols = [_np(m._xyz_motion).reshape(Nd, K * width),
            _np(m._features_dc_motion.transpose(1, 2).flatten(start_dim=1)) if Nd else np.zeros((0, 3), np.float32),
            _np(m._features_rest_motion.transpose(1, 2).flatten(start_dim=1)) if Nd else np.zeros((0, n_rest_m), np.float32),
            _np(m._scaling_motion).reshape(Nd, 3), _np(m._opacity_motion).reshape(Nd, 1),
            _np(m._opacity_duration_center).reshape(Nd, deg), _np(m._opacity_duration_var).reshape(Nd, deg),
            _np(m._rotation_motion).reshape(Nd, K * 4)]
    write_ply(path.replace("point_cloud.ply", "dynamic_point_cloud.ply"),
              dynamic_attributes(K, n_rest_m, width, deg), np.concatenate(cols, axis=1))


def _take(names: List[str], data: np.ndarray, wanted: Sequence[str], path: str) -> np.ndarray:
    pos = {n: i for i, n in enumerate(names)}
    missing = [w for w in wanted if w not in pos]
    if missing:
        raise KeyError("%s: missing PLY properties %s" % (path, missing[:4]))
    idx = [pos[w] for w in wanted]
    if idx == list(range(idx[0], idx[0] + len(idx))):
        return data[:, idx[0]: idx[0] + len(idx)]        # contiguous block: a view, no gather
    return data[:, idx]


def _prefixed(names: List[str], prefix: str, two_level: bool = False) -> List[str]:
    """Property names starting with `prefix`, ordered numerically like load_ply's sorted(..., key=...)
    (by the last index, or by the last two for keyframed attributes)."""
    sel = [n for n in names if n.startswith(prefix)]
    if two_level:
        return sorted(sel, key=lambda x: (int(x.split("_")[-2]), int(x.split("_")[-1])))
    return sorted(sel, key=lambda x: int(x.split("_")[-1]))


def load_model(path: str, sh_degree: int = 3, *, duration: float, interval: float, time_pad: float = 1.0,
               interp_type: str = "cube", rot_interp_type: str = "slerp", var_pad: float = 3.0, kernel_size: float = 0.1,
               device="cpu", check_keyframes: bool = True) -> GaussianArrays:
    """Load …/point_cloud.ply + dynamic_point_cloud.ply (c_gaussian_model.py:558-672).  The scalar
    arguments are the model hyper-parameters the reference takes from its config
    (CGaussianModel.__init__, c_gaussian_model.py:46).  The keyframe count is read off the file; with
    check_keyframes it must equal what load_ply would compute from duration/interval/time_pad."""
    n_rest = 3 * (sh_degree + 1) ** 2 - 3
    names, data = read_ply(path)
    N = data.shape[0]
    rest_names = _prefixed(names, "f_rest_")
    if len(rest_names) != n_rest:
        raise ValueError("%s: %d f_rest_* properties, SH degree %d needs %d" % (path, len(rest_names), sh_degree, n_rest))
    dev = torch.device(device)

    def T(a: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)

    xyz = _take(names, data, ["x", "y", "z"], path)
    f_dc = _take(names, data, ["f_dc_0", "f_dc_1", "f_dc_2"], path).reshape(N, 3, 1)
    f_rest = _take(names, data, rest_names, path).reshape(N, 3, n_rest // 3)
    opac = _take(names, data, ["opacity"], path)
    scales = _take(names, data, _prefixed(names, "scale_"), path)
    rots = _take(names, data, _prefixed(names, "rot_"), path)
    disp = _take(names, data, ["xyz_disp_0", "xyz_disp_1", "xyz_disp_2"], path)

    dpath = path.replace("point_cloud.ply", "dynamic_point_cloud.ply")
    dnames, dd = read_ply(dpath)
    Nd = dd.shape[0]
    xyz_names = _prefixed(dnames, "motion_xyz_", two_level=True)
    rot_names = _prefixed(dnames, "motion_rot_", two_level=True)
    if len(xyz_names) % 3 or len(rot_names) % 4 or len(xyz_names) // 3 != len(rot_names) // 4:
        raise ValueError("%s: inconsistent keyframe columns (%d xyz, %d rot)" % (dpath, len(xyz_names), len(rot_names)))
    K = len(xyz_names) // 3
    shift = _time_shift(time_pad, interval, interp_type)
    dur = max(duration, 1)
    if check_keyframes and Nd and K != expected_keyframes(dur, interval, time_pad, shift):
        raise ValueError("%s holds %d keyframes, the given duration/interval/time_pad imply %d"
                         % (dpath, K, expected_keyframes(dur, interval, time_pad, shift)))
    m_rest_names = _prefixed(dnames, "motion_f_rest_")
    if len(m_rest_names) != n_rest:
        raise ValueError("%s: %d motion_f_rest_* properties, expected %d" % (dpath, len(m_rest_names), n_rest))
    m_xyz = _take(dnames, dd, xyz_names, dpath).reshape(Nd, K, 3)
    m_dc = _take(dnames, dd, ["motion_f_dc_0", "motion_f_dc_1", "motion_f_dc_2"], dpath).reshape(Nd, 3, 1)
    m_rest = _take(dnames, dd, m_rest_names, dpath).reshape(Nd, 3, n_rest // 3)
    m_scale = _take(dnames, dd, _prefixed(dnames, "motion_scale_"), dpath)
    m_opac = _take(dnames, dd, ["motion_opacity"], dpath)
    m_c = _take(dnames, dd, _prefixed(dnames, "motion_opacity_c_"), dpath)
    m_v = _take(dnames, dd, _prefixed(dnames, "motion_opacity_v_"), dpath)
    m_rot = _take(dnames, dd, rot_names, dpath).reshape(Nd, K, 4)

    return GaussianArrays(
        _xyz=T(xyz), _features_dc=T(f_dc.transpose(0, 2, 1)), _features_rest=T(f_rest.transpose(0, 2, 1)),
        _opacity=T(opac), _scaling=T(scales), _rotation=T(rots), _xyz_disp=T(disp),
        _xyz_motion=T(m_xyz), _features_dc_motion=T(m_dc.transpose(0, 2, 1)),
        _features_rest_motion=T(m_rest.transpose(0, 2, 1)), _scaling_motion=T(m_scale), _opacity_motion=T(m_opac),
        _opacity_duration_center=T(m_c.reshape(Nd, -1, 1)), _opacity_duration_var=T(m_v.reshape(Nd, -1, 1)),
        _rotation_motion=T(m_rot),
        max_sh_degree=sh_degree, active_sh_degree=sh_degree, duration=dur, interval=interval, time_pad=time_pad,
        time_shift=shift, var_pad=var_pad, kernel_size=kernel_size, interp_type=interp_type,
        rot_interp_type=rot_interp_type)


# ---------------------------------------------------------------------------------------------
# checkpoint tuple (capture / restore)
# ---------------------------------------------------------------------------------------------
# slot order of CGaussianModel.capture() (c_gaussian_model.py:217-259)
CAPTURE_SLOTS = (
    "active_sh_degree", "_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity",
    "max_radii2D", "min_radii2D", "xyz_gradient_accum", "denom", "xyz_error_accum", "xyz_error_min",
    "xyz_error_min_timestamp", "xyz_ssim_error_accum", "error_denom", "optimizer_state", "spatial_lr_scale",
    "_xyz_disp", "duration", "interval", "time_shift", "keyframe_num", "_xyz_motion", "_features_dc_motion",
    "_features_rest_motion", "_scaling_motion", "_opacity_motion", "_opacity_duration_center", "_opacity_duration_var",
    "_rotation_motion", "motion_max_radii2D", "motion_min_radii2D", "motion_xyz_gradient_accum", "motion_denom",
    "motion_xyz_error_min", "motion_xyz_error_mean", "motion_xyz_error_min_timestamp", "motion_xyz_ssim_error_accum",
    "motion_error_denom",
)


def from_capture(model_args: Sequence, *, time_pad: float = 1.0, interp_type: str = "cube", rot_interp_type: str = "slerp",
                 var_pad: float = 3.0, kernel_size: float = 0.1, sh_degree: int = 3, device=None) -> GaussianArrays:
    """Build the render-path state from the tuple train.py:197 saves as `(gaussians.capture(), iteration)[0]`
    (slot order of c_gaussian_model.py:217-259; restore() :261-320).  Training statistics and the
    optimizer state dict are kept, untouched, in `.extras`."""
    if len(model_args) != len(CAPTURE_SLOTS):
        raise ValueError("capture tuple has %d slots, expected %d" % (len(model_args), len(CAPTURE_SLOTS)))
    d = dict(zip(CAPTURE_SLOTS, model_args))

    def T(x):
        t = x.detach() if isinstance(x, torch.Tensor) else torch.as_tensor(x)
        t = t.to(torch.float32)
        return t.to(device) if device is not None else t

    kw = {n: T(d[n]) for n in STATIC_TENSORS + DYNAMIC_TENSORS}
    extras = {k: v for k, v in d.items() if k not in kw and k not in ("active_sh_degree", "duration", "interval", "time_shift")}
    return GaussianArrays(**kw, max_sh_degree=sh_degree, active_sh_degree=int(d["active_sh_degree"]),
                          duration=d["duration"], interval=d["interval"], time_pad=time_pad, time_shift=d["time_shift"],
                          var_pad=var_pad, kernel_size=kernel_size, interp_type=interp_type,
                          rot_interp_type=rot_interp_type, extras=extras)


def to_capture(m: GaussianArrays) -> tuple:
    """Inverse of from_capture (slots this module does not own come back from `.extras`, else None)."""
    out = []
    for n in CAPTURE_SLOTS:
        if n in STATIC_TENSORS + DYNAMIC_TENSORS or n in ("active_sh_degree", "duration", "interval", "time_shift"):
            out.append(getattr(m, n))
        elif n == "keyframe_num":
            out.append(m.extras.get(n, m.keyframe_num))
        else:
            out.append(m.extras.get(n))
    return tuple(out)


def load_checkpoint(path: str, **kw) -> Tuple[GaussianArrays, int]:
    """chkpnt<iter>.pth as written by train.py:197 -> (arrays, iteration)."""
    model_args, iteration = torch.load(path, map_location="cpu", weights_only=False)
    return from_capture(model_args, **kw), int(iteration)


def load_iteration(model_path: str, iteration: int = -1, **kw) -> Tuple[GaussianArrays, int]:
    """Scene.__init__'s lookup (scene/__init__.py:57-61,173): <model_path>/point_cloud/iteration_<n>/point_cloud.ply,
    n = the largest present when iteration == -1."""
    root = os.path.join(model_path, "point_cloud")
    if iteration == -1:
        its = [int(d.split("_")[-1]) for d in os.listdir(root) if d.startswith("iteration_")]
        if not its:
            raise FileNotFoundError("no iteration_* under %s" % root)
        iteration = max(its)
    p = os.path.join(root, "iteration_%d" % iteration, "point_cloud.ply")
    return load_model(p, **kw), iteration
