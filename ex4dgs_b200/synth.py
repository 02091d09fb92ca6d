"""Seeded synthetic scenes for parity tests and bench.py.

Everything is generated on the CPU in float32 from one ``torch.Generator`` so that the same
bytes reach the oracle, the CUDA path and the compiled reference.  Conventions follow the
reference's camera code (scene/cameras.py:119-127, utils/graphics_utils.py:58-78): matrices are
stored *transposed* (memory is column-major for the device code, auxiliary.h:68-87).

The PyTorch restatement of the model-side getters (pre-interpolation of the dynamic Gaussians for the unchanged-API
path) is oracle material and lives in oracle/getters_oracle.py; the product's own implementation of those getters is
ex4dgs_b200/frontend.py.

Configs (BASELINE.json):
  C1 = 10k static, 400x400            (CPU oracle / CPU baseline)
  C2 = 500k static, 1352x1014, fwd
  C3 = 1.5M static + 0.5M dynamic (K=36 keyframes), 1352x1014, fwd+bwd   <- headline metric
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

SEED = 20240925

CONFIGS = {
    "tiny": dict(P_static=1500, P_dynamic=500, W=160, H=120),
    "C1": dict(P_static=10_000, P_dynamic=0, W=400, H=400),
    "C1d": dict(P_static=7_500, P_dynamic=2_500, W=400, H=400),
    "C2": dict(P_static=500_000, P_dynamic=0, W=1352, H=1014),
    "C3": dict(P_static=1_500_000, P_dynamic=500_000, W=1352, H=1014),
}


@dataclass
class Camera:
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor   # [4,4] world->view, transposed storage
    projmatrix: torch.Tensor   # [4,4] full projection (view @ proj), transposed storage
    campos: torch.Tensor       # [3]
    min_depth: float = 0.2
    max_depth: float = 300.0
    kernel_size: float = 0.1


def projection_matrix(znear: float, zfar: float, tanfovx: float, tanfovy: float) -> torch.Tensor:
    """P as in utils/graphics_utils.py:58-78 (not transposed)."""
    top = tanfovy * znear
    right = tanfovx * znear
    P = torch.zeros(4, 4, dtype=torch.float32)
    P[0, 0] = 2.0 * znear / (2.0 * right)
    P[1, 1] = 2.0 * znear / (2.0 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(W: int, H: int, pose: str = "identity", znear: float = 0.01, zfar: float = 100.0,
                min_depth: float = 0.2, max_depth: float = 300.0, kernel_size: float = 0.1) -> Camera:
    fx = 1462.0 * (W / 1352.0)
    tanfovx = W / (2.0 * fx)
    tanfovy = H / (2.0 * fx)
    if pose == "identity":
        w2c = torch.eye(4, dtype=torch.float32)
    elif pose == "tilted":
        # a fixed, non-trivial rigid transform (exercises every entry of the view matrix)
        ax, ay, az = 0.21, -0.34, 0.13
        cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
        Rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=torch.float64)
        Ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float64)
        Rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=torch.float64)
        w2c = torch.eye(4, dtype=torch.float64)
        w2c[:3, :3] = Rz @ Ry @ Rx
        w2c[:3, 3] = torch.tensor([0.7, -0.4, 1.1], dtype=torch.float64)
        w2c = w2c.float()
    else:
        raise ValueError(pose)
    view_t = w2c.t().contiguous()                                    # world_view_transform
    proj_t = projection_matrix(znear, zfar, tanfovx, tanfovy).t().contiguous()
    full_t = (view_t @ proj_t).contiguous()                          # full_proj_transform
    campos = torch.linalg.inv(view_t)[3, :3].contiguous()
    return Camera(W, H, tanfovx, tanfovy, view_t, full_t, campos, min_depth, max_depth, kernel_size)


@dataclass
class Scene:
    """Raw model tensors (what CGaussianModel holds) + per-frame constants."""
    cam: Camera
    # static Gaussians
    xyz: torch.Tensor            # [Ns,3]
    xyz_disp: torch.Tensor       # [Ns,3]
    rotation: torch.Tensor       # [Ns,4] raw (NOT normalised, c_gaussian_model.py:198)
    scaling: torch.Tensor        # [Ns,3] log-scale
    opacity: torch.Tensor        # [Ns,1] logit
    features: torch.Tensor       # [Ns,16,3]
    # dynamic Gaussians
    xyz_motion: torch.Tensor     # [Nd,K,3]
    rotation_motion: torch.Tensor  # [Nd,K,4]
    scaling_motion: torch.Tensor  # [Nd,3]
    opacity_motion: torch.Tensor  # [Nd,1]
    opacity_center: torch.Tensor  # [Nd,2]
    opacity_var: torch.Tensor     # [Nd,2]
    features_motion: torch.Tensor  # [Nd,16,3]
    # time model
    duration: float = 300.0
    interval: float = 10.0
    time_pad: float = 2.0
    var_pad: float = 3.0
    timestamp: float = 137.0
    sh_degree: int = 3
    bg: torch.Tensor = field(default_factory=lambda: torch.ones(3))

    @property
    def time_shift(self) -> float:      # "cube": time_pad + interval (c_gaussian_model.py:76,119)
        return self.time_pad + self.interval

    @property
    def P(self) -> int:
        return self.xyz.shape[0] + self.xyz_motion.shape[0]


def _camera_space_points(n, cam: Camera, g):
    z = torch.exp(torch.empty(n).uniform_(math.log(1.5), math.log(60.0), generator=g))
    ndc = torch.empty(n, 2).uniform_(-1.4, 1.4, generator=g)
    p = torch.stack([ndc[:, 0] * cam.tanfovx * z, ndc[:, 1] * cam.tanfovy * z, z], dim=1)
    return p, z


def make_scene(P_static: int, P_dynamic: int, W: int, H: int, seed: int = SEED, pose: str = "identity",
               K: int = 36, bg: Optional[torch.Tensor] = None, dir_nonzero: bool = False,
               sigma_px: float = 2.0, **cam_kw) -> Scene:
    g = torch.Generator().manual_seed(seed)
    cam = make_camera(W, H, pose, **cam_kw)
    fx = W / (2.0 * cam.tanfovx)
    n = P_static + P_dynamic
    p_cam, z = _camera_space_points(n, cam, g)
    # camera -> world
    c2w = torch.linalg.inv(cam.viewmatrix.t())
    p_world = p_cam @ c2w[:3, :3].t() + c2w[:3, 3]
    sig = torch.exp(math.log(sigma_px) + 0.7 * torch.randn(n, generator=g))
    aniso = torch.exp(0.5 * torch.randn(n, 3, generator=g))
    scale = sig[:, None] * aniso * z[:, None] / fx
    q = torch.randn(n, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    q = q * (1.0 + 1e-3 * torch.randn(n, 1, generator=g))
    opac_logit = 2.0 * torch.randn(n, 1, generator=g)
    feats = torch.cat([torch.randn(n, 1, 3, generator=g), 0.15 * torch.randn(n, 15, 3, generator=g)], dim=1)

    ns = P_static
    sc = dict(cam=cam,
              xyz=p_world[:ns].contiguous(),
              xyz_disp=(0.01 * z[:ns, None] * torch.randn(ns, 3, generator=g)).contiguous(),
              rotation=q[:ns].contiguous(), scaling=torch.log(scale[:ns]).contiguous(),
              opacity=opac_logit[:ns].contiguous(), features=feats[:ns].contiguous())
    nd = P_dynamic
    steps = 0.02 * z[ns:, None, None] * torch.randn(nd, K, 3, generator=g)
    # random walk centred on the base mean at the keyframe used by the bench timestamp
    walk = torch.cumsum(steps, dim=1)
    walk = walk - walk[:, K // 2 - 3:K // 2 - 2, :]
    sc.update(xyz_motion=(p_world[ns:, None, :] + walk).contiguous(),
              rotation_motion=(q[ns:, None, :] + 0.05 * torch.randn(nd, K, 4, generator=g)).contiguous(),
              scaling_motion=torch.log(scale[ns:]).contiguous(),
              opacity_motion=opac_logit[ns:].contiguous(),
              features_motion=feats[ns:].contiguous())
    tau_max = (300.0 + 12.0) / 10.0
    c = torch.sort(torch.empty(nd, 2).uniform_(0.0, tau_max, generator=g), dim=1)[0]
    sc.update(opacity_center=c.contiguous(),
              opacity_var=(1.0 + 0.5 * torch.randn(nd, 2, generator=g)).contiguous())
    scene = Scene(**sc)
    if bg is not None:
        scene.bg = bg.float()
    scene._dir_nonzero = dir_nonzero
    scene._seed = seed
    return scene


def make_config(name: str, **kw) -> Scene:
    cfg = dict(CONFIGS[name])
    cfg.update(kw)
    return make_scene(**cfg)


def grad_outputs(sc: Scene, seed_offset: int = 7) -> Dict[str, torch.Tensor]:
    """Upstream gradients as in SURVEY 8d / train.py:148-153: dense colour grad, and the
    [acc, l1, ssim]-style non-negative tensor routed in as grad_flow."""
    H, W = sc.cam.H, sc.cam.W
    g = torch.Generator().manual_seed(getattr(sc, "_seed", SEED) + seed_offset)
    gc = (torch.rand(3, H, W, generator=g) * 2 - 1) / (3 * H * W)
    gf = torch.rand(3, H, W, generator=g)
    return dict(grad_color=gc, grad_depth=torch.zeros(1, H, W), grad_flow=gf, grad_acc=torch.zeros(1, H, W))
