#!/usr/bin/env python
"""bench.py - headline metric of BASELINE.json: rasterizer fwd+bwd frames/s at config C3
(2.0M static+dynamic Gaussians, 1352x1014), plus the render-forward roofline figure.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" = one frame: GaussianRasterizer forward + backward on the pre-interpolated [P,.] tensors
(the drop-in boundary).  With N GPUs every rank renders its own frame per step
(frames shard one-per-GPU, Gaussians replicated: weak scaling), the only collective is one NCCL
all-reduce of the scalar loss in the end-to-end leg.  Prints ONE JSON line (rank 0).

  value   frames/s, inputs resident in HBM, upstream gradients resident (device-timed, max over ranks)
  e2e     frames/s through the public API with the per-frame HOST inputs (camera + ground-truth image,
          pinned) copied H2D inside the timed region, L1 loss, backward, loss copied D2H.  The
          Gaussian parameters are model state and stay resident (as weights do in a training step).
  roofline  render-forward kernel: algorithmic bytes (56*R_eff + 52*W*H + 8*tiles) / its
          CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the CPU oracle (oracle/cpu_raster.c, a port of the reference algorithm: the reference
          has no CPU path) on one full frame of the same workload, all host threads.

--impl reference runs the UNMODIFIED reference CUDA extension (oracle/_ref, built by
oracle/build_ref.py) on the same GPU, same workload and step definition; if that build is absent it
falls back to the CPU oracle port (see DESIGN.md "reference arm").
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from ex4dgs_b200 import synth  # noqa: E402

METRIC = "fwd+bwd frames/sec @2.0M Gaussians 1352x1014"
FALLBACK_HBM_GBS = 6650.0


def dist_env():
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), ws


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return None
        time.sleep(0.12)
        self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.06 or ts > t1 + 0.06:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                smax = max(smax, float(f[2]))
            except Exception:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": smax or None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def load_reference():
    so = os.path.join(ROOT, "oracle", "_ref", "diff_gaussian_rasterization_df", "_C.so")
    if not os.path.exists(so):
        return None
    import importlib.util
    d = os.path.dirname(so)
    spec = importlib.util.spec_from_file_location("ref_diff_gaussian_rasterization_df", os.path.join(d, "__init__.py"),
                                                  submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_diff_gaussian_rasterization_df"] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_loss():
    """utils/loss_utils.py of the reference, installed unmodified next to the extension (oracle/build_ref.py)."""
    d = os.path.join(ROOT, "oracle", "_ref", "callers")
    if not os.path.exists(os.path.join(d, "utils", "loss_utils.py")):
        return None
    sys.path.insert(0, d)
    from utils import loss_utils
    return loss_utils


def load_reference_model_class():
    """The reference's CGaussianModel class (scene/c_gaussian_model.py, installed unmodified by oracle/build_ref.py):
    its statistics methods are what the reference arm's training iteration runs.  Two of its imports do not exist
    in this image and are never used by those methods: `plyfile` and `simple_knn._C` get empty stand-ins."""
    d = os.path.join(ROOT, "oracle", "_ref", "callers")
    path = os.path.join(d, "scene", "c_gaussian_model.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    import types
    if "plyfile" not in sys.modules:
        m = types.ModuleType("plyfile")
        m.PlyData = m.PlyElement = object
        sys.modules["plyfile"] = m
    if "simple_knn._C" not in sys.modules:
        pkg, sub = types.ModuleType("simple_knn"), types.ModuleType("simple_knn._C")
        sub.distCUDA2 = None
        pkg._C = sub
        sys.modules["simple_knn"], sys.modules["simple_knn._C"] = pkg, sub
    if d not in sys.path:
        sys.path.insert(0, d)
    spec = importlib.util.spec_from_file_location("ref_c_gaussian_model", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CGaussianModel


def stats_tensors(Ns, Nd, dev):
    """CGaussianModel.training_setup (scene/c_gaussian_model.py:412-428) and :408-409 / :843-844."""
    z = lambda *s: torch.zeros(*s, device=dev)
    o = lambda *s: torch.ones(*s, device=dev)
    return dict(
        max_radii2D=z(Ns), min_radii2D=o(Ns) * 1000, xyz_gradient_accum=z(Ns, 1), denom=z(Ns, 1), xyz_error_accum=z(Ns, 1),
        xyz_error_min=o(Ns, 1) * 1000, xyz_error_min_timestamp=o(Ns, 1) * -1, xyz_ssim_error_accum=z(Ns, 1), error_denom=z(Ns, 1),
        motion_max_radii2D=z(Nd), motion_min_radii2D=o(Nd) * 1000, motion_xyz_gradient_accum=z(Nd, 1), motion_denom=z(Nd, 1),
        motion_xyz_error_min=o(Nd, 1) * 1000, motion_xyz_error_mean=z(Nd, 1), motion_xyz_error_min_timestamp=o(Nd, 1) * -1,
        motion_xyz_ssim_error_accum=z(Nd, 1), motion_error_denom=z(Nd, 1))


def reference_model(sc: synth.Scene, dev, cls):
    """The reference's UNMODIFIED CGaussianModel (oracle/_ref/callers/scene/c_gaussian_model.py) built by its own
    constructor and holding the synthetic scene in nn.Parameters of the reference's shapes."""
    prev = torch.cuda.current_device()
    torch.cuda.set_device(dev)          # the class calls .cuda() without a device
    try:
        m = cls(sh_degree=3, duration=int(sc.duration), interval=int(sc.interval), time_pad=int(sc.time_pad),
                interp_type="cube", rot_interp_type="slerp", var_pad=sc.var_pad, kernel_size=sc.cam.kernel_size)
    finally:
        torch.cuda.set_device(prev)
    P = lambda t: torch.nn.Parameter(t.detach().to(dev).float().contiguous().clone())     # noqa: E731
    m._xyz, m._xyz_disp, m._rotation, m._scaling, m._opacity = P(sc.xyz), P(sc.xyz_disp), P(sc.rotation), P(sc.scaling), P(sc.opacity)
    m._features_dc, m._features_rest = P(sc.features[:, :1]), P(sc.features[:, 1:])
    m._xyz_motion, m._rotation_motion = P(sc.xyz_motion), P(sc.rotation_motion)
    m._scaling_motion, m._opacity_motion = P(sc.scaling_motion), P(sc.opacity_motion)
    m._opacity_duration_center, m._opacity_duration_var = P(sc.opacity_center[:, :, None]), P(sc.opacity_var[:, :, None])
    m._features_dc_motion, m._features_rest_motion = P(sc.features_motion[:, :1]), P(sc.features_motion[:, 1:])
    m.active_sh_degree = sc.sh_degree
    return m


def scene_model(sc: synth.Scene, dev):
    """The same scene as a plain namespace with the reference's attribute names (what FusedGetters wraps)."""
    from types import SimpleNamespace
    P = lambda t: torch.nn.Parameter(t.detach().to(dev).float().contiguous().clone())     # noqa: E731
    return SimpleNamespace(
        _xyz=P(sc.xyz), _xyz_disp=P(sc.xyz_disp), _rotation=P(sc.rotation), _scaling=P(sc.scaling), _opacity=P(sc.opacity),
        _xyz_motion=P(sc.xyz_motion), _rotation_motion=P(sc.rotation_motion), _scaling_motion=P(sc.scaling_motion),
        _opacity_motion=P(sc.opacity_motion), _opacity_duration_center=P(sc.opacity_center[:, :, None]),
        _opacity_duration_var=P(sc.opacity_var[:, :, None]), _features_dc=P(sc.features[:, :1]),
        _features_rest=P(sc.features[:, 1:]), _features_dc_motion=P(sc.features_motion[:, :1]),
        _features_rest_motion=P(sc.features_motion[:, 1:]),
        duration=sc.duration, interval=sc.interval, time_shift=sc.time_shift, var_pad=sc.var_pad,
        kernel_size=sc.cam.kernel_size, active_sh_degree=sc.sh_degree, max_sh_degree=3)


def scene_inputs(sc: synth.Scene, dev, impl: str):
    """The flat [P, .] tensors gaussian_renderer/__init__.py:62-95 hands to the rasterizer (static first, then dynamic,
    pre-interpolated at the scene's timestamp), produced - outside every timed region - by the arm's OWN getters:
    this repository's fused front-end for our arm, the reference's CGaussianModel getters for the reference arm."""
    t = sc.timestamp
    with torch.no_grad():
        if impl == "ours":
            from ex4dgs_b200.frontend import FusedGetters
            g = FusedGetters(scene_model(sc, dev))
            means, rots, scales, opac = g.get_xyz_at_t(t), g.get_rotation_at_t(t), g.get_scaling(), g.get_opacity_at_t(t)
            shs = g.get_features().cat()
        else:
            cls = load_reference_model_class()
            if cls is None:
                raise SystemExit("reference arm: oracle/_ref/callers/scene/c_gaussian_model.py is missing (run oracle/build_ref.py)")
            g = reference_model(sc, dev, cls)
            means, rots, scales, opac = g.get_xyz_at_t(t), g.get_rotation_at_t(t), g.get_scaling(), g.get_opacity_at_t(t)
            shs = g.get_features()
    return dict(means3D=means.float().contiguous(), dir3D=torch.zeros_like(means), opacities=opac.float().contiguous(),
                shs=shs.float().contiguous(), scales=scales.float().contiguous(), rotations=rots.float().contiguous())



def kernel_source_hash():
    """sha256/16 of the compositing kernels' sources: profiles/traffic.json records it at capture time, so that the
    per-launch figures taken from the committed ncu capture (DRAM bytes, instruction counts) are only used for the code
    they were measured on."""
    import hashlib
    h = hashlib.sha256()
    for f in ("render_fwd.cu", "render_bwd.cu", "common.cuh"):
        with open(os.path.join(ROOT, "ex4dgs_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def load_reference_render():
    """gaussian_renderer/__init__.py of the reference (installed unmodified under oracle/_ref/callers by
    oracle/build_ref.py) - the ONE caller of the rasterizer; it imports `diff_gaussian_rasterization_df` by name."""
    d = os.path.join(ROOT, "oracle", "_ref", "callers")
    if not os.path.exists(os.path.join(d, "gaussian_renderer", "__init__.py")):
        return None
    import importlib
    if d not in sys.path:
        sys.path.insert(0, d)
    for m in ("gaussian_renderer",):
        sys.modules.pop(m, None)
    return importlib.import_module("gaussian_renderer")


def make_render_api_step(frame: "Frame", impl: str, raster_mod):
    """the frame through the UNCHANGED gaussian_renderer.render() (per-frame getters, torch.zeros_like
    screen-space tensors, the torch.cuda.synchronize() at its end) + the same backward.  Our arm hands render() a
    FusedGetters wrapper around the model and this repository's drop-in package; the reference arm its own
    CGaussianModel and extension."""
    import math
    import types
    sys.modules["diff_gaussian_rasterization_df"] = raster_mod if impl != "ours" else __import__("diff_gaussian_rasterization_df")
    gr = load_reference_render()
    if gr is None:
        return None
    sc, dev = frame.sc, frame.dev
    cam = sc.cam
    if impl == "ours":
        from ex4dgs_b200.frontend import FusedGetters
        pc = FusedGetters(scene_model(sc, dev))
    else:
        cls = load_reference_model_class()
        if cls is None:
            return None
        pc = reference_model(sc, dev, cls)
    vc = types.SimpleNamespace(FoVx=2 * math.atan(cam.tanfovx), FoVy=2 * math.atan(cam.tanfovy), image_height=cam.H,
                               image_width=cam.W, world_view_transform=frame.view, full_proj_transform=frame.proj,
                               camera_center=frame.campos, timestamp=sc.timestamp)
    pipe = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    params = [getattr(pc, n) for n in ("_xyz", "_xyz_disp", "_rotation", "_scaling", "_opacity", "_xyz_motion", "_rotation_motion",
                                       "_scaling_motion", "_opacity_motion", "_opacity_duration_center", "_opacity_duration_var",
                                       "_features_dc", "_features_rest", "_features_dc_motion", "_features_rest_motion")]

    def step():
        out = gr.render(vc, pc, pipe, frame.bg, near=cam.min_depth, far=cam.max_depth, timestamp=sc.timestamp)
        torch.autograd.backward([out["render"], out["opticalflow"]], [frame.go["grad_color"], frame.go["grad_flow"]])
        for q in params:
            q.grad = None
    return step


class Frame:
    """Resident inputs of one rank + the step functions."""

    def __init__(self, mod, sc: synth.Scene, dev, seed_off: int, impl: str = "ours"):
        self.mod, self.sc, self.dev, self.impl = mod, sc, dev, impl
        cam = sc.cam
        inp = scene_inputs(sc, dev, impl)
        self.P = inp["means3D"].shape[0]
        self.t = {k: v.detach().clone().requires_grad_(True) for k, v in inp.items()}
        self.means2D = torch.zeros(self.P, 3, device=dev, requires_grad=True)
        go = synth.grad_outputs(sc, seed_offset=7 + seed_off)
        self.go = {k: v.to(dev) for k, v in go.items()}
        self.sub = torch.zeros(cam.H, cam.W, 2, device=dev)
        self.bg = sc.bg.to(dev)
        # resident camera (device-timed leg) and pinned host copies (end-to-end leg)
        self.view, self.proj, self.campos = cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.campos.to(dev)
        g = torch.Generator().manual_seed(1234 + seed_off)
        self.h_gt = torch.rand(3, cam.H, cam.W, generator=g).pin_memory()
        self.h_cam = torch.cat([cam.viewmatrix.flatten(), cam.projmatrix.flatten(), cam.campos.flatten()]).pin_memory()
        # end-to-end legs: the ground-truth image of step i+1 is uploaded (from pinned memory, on the copy stream) while
        # step i computes - one image per step inside the timed region, two device buffers; the loss all-reduce and its
        # D2H run beside the next step on their own stream (ring of 4 loss words), so that no rank's compute stream waits
        # for the slowest rank inside a step
        self.d_gt2 = [torch.empty(3, cam.H, cam.W, device=dev) for _ in range(2)]
        self.gt_ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.gt_free = [torch.cuda.Event(), torch.cuda.Event()]
        self.gt_step = 0
        self.d_gt = self.d_gt2[0]
        self.d_cam = torch.empty(35, device=dev)
        self.h_loss = torch.zeros(1).pin_memory()
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.loss_stream = torch.cuda.Stream(device=dev)
        self.loss_ring = [torch.zeros(1, device=dev) for _ in range(4)]
        self.loss_done = [torch.cuda.Event() for _ in range(4)]
        self.loss_step = 0
        self.R = -1
        self.last = None
        # the L1 loss of the e2e step: utils/loss_utils.py:22-25 (torch ops) for the reference arm, this repository's
        # drop-in for it (ex4dgs_b200/loss.py::l1_loss, one kernel each way) for our arm
        self.l1 = None
        if getattr(mod, "__name__", "") == "ex4dgs_b200":
            from ex4dgs_b200.loss import l1_loss
            self.l1 = l1_loss

    def settings(self, view, proj, campos):
        cam = self.sc.cam
        return self.mod.GaussianRasterizationSettings(
            image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, kernel_size=cam.kernel_size,
            subpixel_offset=self.sub, bg=self.bg, scale_modifier=1.0, viewmatrix=view, projmatrix=proj,
            sh_degree=self.sc.sh_degree, campos=campos, prefiltered=False, min_depth=cam.min_depth,
            max_depth=cam.max_depth, debug=False)

    def _raster(self, rs):
        t = self.t
        return self.mod.GaussianRasterizer(rs)(means3D=t["means3D"], means2D=self.means2D, dir3D=t["dir3D"],
                                                opacities=t["opacities"], shs=t["shs"], scales=t["scales"],
                                                rotations=t["rotations"])

    def _zero(self):
        for v in list(self.t.values()) + [self.means2D]:
            v.grad = None

    def step_forward(self):
        with torch.no_grad():
            self.last = self._raster(self.settings(self.view, self.proj, self.campos))[0]

    def step_device(self):
        color, radii, depth, flow, acc, idxs = self._raster(self.settings(self.view, self.proj, self.campos))
        go = self.go
        # dense grad_color, the hook tensor as grad_flow, grad_depth = grad_acc = 0.  As in train.py
        # the depth / acc images simply do not enter the loss: autograd hands the reference materialised zero
        # tensors for them (its Function keeps the default), this repo's Function receives None -> NULL.
        torch.autograd.backward([color, flow], [go["grad_color"], go["grad_flow"]])
        self.last = color
        if color.grad_fn is not None and hasattr(color.grad_fn, "num_rendered"):
            self.R = int(color.grad_fn.num_rendered)
        self._zero()

    def gt_begin(self):
        """H2D of this step's camera (main stream) and of the NEXT step's ground truth (copy stream, into the buffer the
        previous step's loss has released); returns the device image of THIS step once its upload is awaited."""
        main = torch.cuda.current_stream(self.dev)
        self.d_cam.copy_(self.h_cam, non_blocking=True)
        i = self.gt_step
        self.gt_step += 1
        if i == 0:                                   # very first step: nothing was prefetched yet
            with torch.cuda.stream(self.copy_stream):
                self.d_gt2[0].copy_(self.h_gt, non_blocking=True)
                self.gt_ready[0].record(self.copy_stream)
        nxt = (i + 1) & 1
        with torch.cuda.stream(self.copy_stream):
            if i >= 1:
                self.copy_stream.wait_event(self.gt_free[nxt])      # loss of step i-1 has read that buffer
            self.d_gt2[nxt].copy_(self.h_gt, non_blocking=True)
            self.gt_ready[nxt].record(self.copy_stream)
        return i & 1

    def gt_wait(self, buf):
        main = torch.cuda.current_stream(self.dev)
        main.wait_event(self.gt_ready[buf])
        self.d_gt = self.d_gt2[buf]
        return self.d_gt

    def gt_release(self, buf):
        self.gt_free[buf].record(torch.cuda.current_stream(self.dev))

    def loss_out(self, loss, group):
        """all-reduce of the scalar loss (the ONE collective of the path) + D2H, off the compute stream"""
        main = torch.cuda.current_stream(self.dev)
        k = self.loss_step & 3
        self.loss_step += 1
        main.wait_event(self.loss_done[k])           # the ring slot's previous use (4 steps ago) is over
        buf = self.loss_ring[k]
        buf.copy_(loss.detach().reshape(1))
        self.loss_stream.wait_stream(main)
        with torch.cuda.stream(self.loss_stream):
            if group is not None:
                torch.distributed.all_reduce(buf, group=group)
            self.h_loss.copy_(buf, non_blocking=True)
            self.loss_done[k].record(self.loss_stream)

    def step_e2e(self, group=None):
        buf = self.gt_begin()
        c = self.d_cam
        rs = self.settings(c[0:16].view(4, 4), c[16:32].view(4, 4), c[32:35])
        color, radii, depth, flow, acc, idxs = self._raster(rs)
        gt = self.gt_wait(buf)
        loss = self.l1(color, gt) if self.l1 is not None else torch.abs((color - gt)).mean()
        self.gt_release(buf)
        torch.autograd.backward([loss, flow], [None, self.go["grad_flow"]])
        self.loss_out(loss, group)
        self._zero()


def make_train_step(frame: Frame, ref_loss, lambda_dssim=0.2):
    """The training iteration of train.py:139-172 around the rasterizer: H2D camera + ground truth,
    render, loss = 0.8 L1 + 0.2 (1 - SSIM), the l1_accum hook tensor as the flow gradient, backward,
    loss D2H.  ref_loss = the reference's utils/loss_utils module (reference arm) or None (fused loss)."""
    if ref_loss is None:
        from ex4dgs_b200.loss import photometric_loss

    def step(group=None):
        buf = frame.gt_begin()
        c = frame.d_cam
        image, radii, depth, flow, acc, idxs = frame._raster(frame.settings(c[0:16].view(4, 4), c[16:32].view(4, 4), c[32:35]))
        gt_image = frame.gt_wait(buf)
        if ref_loss is None:
            loss, Ll1, _, l1_errors, ssim_errors = photometric_loss(image, gt_image, lambda_dssim)
        else:
            Ll1 = ref_loss.l1_loss(image, gt_image)
            loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - ref_loss.ssim(image, gt_image))
            l1_errors = (image - gt_image).abs().mean(dim=0)
            ssim_errors = ref_loss.ssim(image, gt_image, reduce=False).mean(dim=0)
        frame.gt_release(buf)
        hook_tensor = torch.stack([acc[0], l1_errors, ssim_errors])
        flow_h = flow.register_hook(lambda grad: hook_tensor)
        loss = loss + flow.mean() * 0
        loss.backward()
        flow_h.remove()                                  # train.py:173
        frame.loss_out(loss, group)
        frame._zero()
    return step

# learning rates of arguments/__init__.py:93-110 (OptimizationParams defaults), in the group order of
# scene/c_gaussian_model.py:430-449
MODEL_GROUPS = [("xyz", 1.6e-4), ("f_dc", 2.5e-3), ("f_rest", 2.5e-3 / 20), ("opacity", 0.05), ("scaling", 0.005),
                ("rotation", 1e-5), ("xyz_disp", 1e-4), ("motion_xyz", 1.6e-4), ("motion_f_dc", 2.5e-3),
                ("motion_f_rest", 2.5e-3 / 20), ("motion_scaling", 0.005), ("motion_opacity", 0.05),
                ("motion_opacity_center", 1e-3), ("motion_opacity_var", 5e-4), ("motion_rotation", 1e-3)]
LR_SCALE = 1e-3     # keeps the synthetic scene stationary over the timed steps; RAdam's work is lr-independent


DP_OVERLAP = int(os.environ.get("EX4DGS_DP_OVERLAP", "1"))    # --dp-grads: 1 = all-reduce overlapped with the bucketed RAdam step
STATIC_REG, MOTION_REG = 0.0001, 0.0001      # arguments/__init__.py:134-135 (rot_reg = 0.0: its branch never runs)


def make_model_step(frame: Frame, impl: str, ref_loss, dp_grads: bool, lambda_dssim=0.2, bookkeeping=True, ref_model_cls=None):
    """One full training iteration on the MODEL's native parameters (train.py:124-253 minus
    densify_and_prune): per-frame getters (fused front-end kernel, row N1 / the getters of the reference's own
    CGaussianModel instance, scene/c_gaussian_model.py:170-215,330-375), get_features' torch.cat, render,
    loss block (row N2 / utils/loss_utils.py), backward down to the 15 parameter tensors,
    optimizer.step() (FusedRAdam, row N4 / torch.optim.RAdam) and zero_grad(set_to_none=True).
    bookkeeping adds what train.py does around that every iteration: the static / motion regularisation terms
    (:156-162), mark_prune_stats + max_radii2D + add_densification_stats + add_l1_ssim_stats (:196-212), the
    nan_to_num of one gradient (:246-248) and prune_nan_points' NaN test (:253) - fused kernels
    (ex4dgs_b200/stats.py, FusedRAdam guards) against the reference's own methods and expressions."""
    import copy
    sc, dev = frame.sc, frame.dev
    Ns, Nd = sc.xyz.shape[0], sc.xyz_motion.shape[0]
    if impl == "ours":
        m = copy.copy(sc)

        def par(t):
            return torch.nn.Parameter(t.detach().to(dev).float().contiguous().clone())

        raw = ["xyz", "xyz_disp", "rotation", "scaling", "opacity", "xyz_motion", "rotation_motion", "scaling_motion",
               "opacity_motion", "opacity_center", "opacity_var"]
        for n in raw:
            setattr(m, n, par(getattr(sc, n)))
        f_dc, f_rest = par(sc.features[:, :1]), par(sc.features[:, 1:])
        f_dc_m, f_rest_m = par(sc.features_motion[:, :1]), par(sc.features_motion[:, 1:])
        by_name = {"xyz": m.xyz, "f_dc": f_dc, "f_rest": f_rest, "opacity": m.opacity, "scaling": m.scaling,
                   "rotation": m.rotation, "xyz_disp": m.xyz_disp, "motion_xyz": m.xyz_motion, "motion_f_dc": f_dc_m,
                   "motion_f_rest": f_rest_m, "motion_scaling": m.scaling_motion, "motion_opacity": m.opacity_motion,
                   "motion_opacity_center": m.opacity_center, "motion_opacity_var": m.opacity_var,
                   "motion_rotation": m.rotation_motion}
    else:
        # the reference arm runs the reference's OWN model class: its getters (get_xyz_at_t, get_opacity_at_t, get_scaling,
        # get_rotation_at_t, get_features - what gaussian_renderer/__init__.py:62-95 calls) and its statistics methods
        if ref_model_cls is None:
            ref_model_cls = load_reference_model_class()
        if ref_model_cls is None:
            raise SystemExit("reference arm: oracle/_ref/callers/scene/c_gaussian_model.py is missing (run oracle/build_ref.py)")
        gm = reference_model(sc, dev, ref_model_cls)
        by_name = {"xyz": gm._xyz, "f_dc": gm._features_dc, "f_rest": gm._features_rest, "opacity": gm._opacity,
                   "scaling": gm._scaling, "rotation": gm._rotation, "xyz_disp": gm._xyz_disp, "motion_xyz": gm._xyz_motion,
                   "motion_f_dc": gm._features_dc_motion, "motion_f_rest": gm._features_rest_motion,
                   "motion_scaling": gm._scaling_motion, "motion_opacity": gm._opacity_motion,
                   "motion_opacity_center": gm._opacity_duration_center, "motion_opacity_var": gm._opacity_duration_var,
                   "motion_rotation": gm._rotation_motion}
    groups = [{"params": [by_name[n]], "lr": lr * LR_SCALE, "name": n} for n, lr in MODEL_GROUPS]
    params = [g["params"][0] for g in groups]
    if impl == "ours":
        from types import SimpleNamespace
        from ex4dgs_b200 import stats as fstats
        from ex4dgs_b200.frontend import interpolate_gaussians
        from ex4dgs_b200.loss import photometric_loss
        from ex4dgs_b200.optim import FusedRAdam, allreduce_gradients
        from ex4dgs_b200.rasterizer import SegmentedSH
        if bookkeeping:
            opt = FusedRAdam(groups, lr=0.001, check_nan=("xyz", "motion_xyz"), sanitize_grad=("motion_opacity_var",))
        else:
            opt = FusedRAdam(groups, lr=0.001)
        gaussians = SimpleNamespace(_xyz=m.xyz, **stats_tensors(Ns, Nd, dev))
    else:
        opt = torch.optim.RAdam(groups, lr=0.001)
        gaussians = gm
        for k, v in stats_tensors(Ns, Nd, dev).items():
            setattr(gaussians, k, v)
    P = Ns + Nd
    n_param = sum(p.numel() for p in params)

    def step(group=None):
        buf = frame.gt_begin()
        c = frame.d_cam
        rs = frame.settings(c[0:16].view(4, 4), c[16:32].view(4, 4), c[32:35])
        if impl == "ours":
            shs = SegmentedSH(f_dc, f_rest, f_dc_m, f_rest_m)       # the four tensors read in place, no cat
            means, rots, scales, opac = interpolate_gaussians(
                m.xyz, m.xyz_disp, m.rotation, m.scaling, m.opacity, m.xyz_motion, m.rotation_motion, m.scaling_motion,
                m.opacity_motion, m.opacity_center, m.opacity_var, t=sc.timestamp, duration=sc.duration,
                interval=sc.interval, time_shift=sc.time_shift, var_min=sc.var_pad / sc.interval)
        else:
            # gaussian_renderer/__init__.py:62-95 on the reference's own class
            means = gaussians.get_xyz_at_t(sc.timestamp, mode=0, training=True)
            opac = gaussians.get_opacity_at_t(sc.timestamp, mode=0, training=True)
            scales = gaussians.get_scaling(mode=0)
            rots = gaussians.get_rotation_at_t(sc.timestamp, mode=0)
            shs = gaussians.get_features(mode=0)
        means2D = torch.zeros(P, 3, device=dev, requires_grad=True)     # gaussian_renderer/__init__.py:28
        flow_in = torch.zeros(P, 3, device=dev, requires_grad=True)      # :66
        image, radii, depth, flow, acc, idxs = frame.mod.GaussianRasterizer(rs)(
            means3D=means, means2D=means2D, dir3D=flow_in, opacities=opac, shs=shs, scales=scales, rotations=rots)
        gt_image = frame.gt_wait(buf)
        if impl == "ours":
            loss, Ll1, _, l1_errors, ssim_errors = photometric_loss(image, gt_image, lambda_dssim)
        else:
            Ll1 = ref_loss.l1_loss(image, gt_image)
            loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - ref_loss.ssim(image, gt_image))
            l1_errors = (image - gt_image).abs().mean(dim=0)
            ssim_errors = ref_loss.ssim(image, gt_image, reduce=False).mean(dim=0)
        frame.gt_release(buf)
        hook_tensor = torch.stack([acc[0], l1_errors, ssim_errors])
        flow_h = flow.register_hook(lambda grad: hook_tensor)
        loss = loss + flow.mean() * 0
        if bookkeeping and impl != "ours":
            # train.py:156-162, verbatim
            loss += STATIC_REG * torch.log(gaussians._xyz_disp.norm(dim=-1) + 0.001).mean()
            diff1 = (gaussians._xyz_motion[:, :1] - gaussians._xyz_motion[:, 1:])
            loss += MOTION_REG * diff1.norm(dim=-1).mean()
        loss.backward()
        flow_h.remove()
        if bookkeeping and impl == "ours":
            # same two terms: values added to the reported loss, gradients added to the .grad of the two tensors
            loss = loss.detach() + fstats.regularizers_(m.xyz_disp, m.xyz_motion, STATIC_REG, MOTION_REG).sum()
        if bookkeeping:
            with torch.no_grad():
                if impl == "ours":
                    fstats.iteration_stats(gaussians, radii, means2D.grad, flow_in.grad, sc.timestamp, densify=True, static_num=Ns)
                else:
                    # train.py:196-212, verbatim (viewspace_point_tensor = means2D, viewspace_point_error_tensor = flow_in)
                    viewspace_point_tensor, viewspace_point_error_tensor, visibility_filter = means2D, flow_in, radii > 0
                    gaussians.mark_prune_stats(radii, viewspace_point_error_tensor)
                    static_num = gaussians._xyz.shape[0]
                    static_vis_filter = visibility_filter[:static_num]
                    static_radii = radii[:static_num]
                    dynamic_vis_filter = visibility_filter[static_num:]
                    dynamic_radii = radii[static_num:]
                    gaussians.max_radii2D[static_vis_filter] = torch.max(gaussians.max_radii2D[static_vis_filter], static_radii[static_vis_filter])
                    gaussians.motion_max_radii2D[dynamic_vis_filter] = torch.max(gaussians.motion_max_radii2D[dynamic_vis_filter], dynamic_radii[dynamic_vis_filter])
                    gaussians.add_densification_stats(viewspace_point_tensor, static_vis_filter, dynamic_vis_filter, static_num)
                    gaussians.add_l1_ssim_stats(viewspace_point_error_tensor, static_vis_filter, dynamic_vis_filter, static_num, sc.timestamp)
                    # train.py:244-248
                    if gaussians._opacity_duration_var.shape[0] != 0:
                        if not gaussians._opacity_duration_var.grad is None:
                            gaussians._opacity_duration_var.grad = gaussians._opacity_duration_var.grad.nan_to_num()
        if impl == "ours":
            if dp_grads and group is not None and DP_OVERLAP:
                opt.step(allreduce_group=group)        # reductions queued up front, the step kernel follows them bucket by bucket
            elif dp_grads and group is not None:
                opt.step(grad_scale=allreduce_gradients(params, group))
            else:
                opt.step()
        else:
            if dp_grads and group is not None:
                ws = torch.distributed.get_world_size(group)
                for p in params:
                    torch.distributed.all_reduce(p.grad, group=group)
                    p.grad /= ws
            opt.step()
        opt.zero_grad(set_to_none=True)
        if bookkeeping:
            if impl == "ours":
                if any(opt.poll_nan().values()):                  # flags raised by the step kernel, read without waiting
                    raise RuntimeError("NaN parameters")
            else:
                with torch.no_grad():
                    gaussians.prune_nan_points()                  # train.py:253
        frame.loss_out(loss, group)
    # handles for tests/test_gpu_train_iter.py (trajectory parity of the two arms)
    step.params = dict(zip([n for n, _ in MODEL_GROUPS], params))
    step.gaussians = gaussians
    step.optimizer = opt
    return step, n_param


def frame_stats(frame: Frame):
    """R, P_vis and R_eff = sum_tiles min(range_len, batch size * batches fetched) from the scratch buffers."""
    from ex4dgs_b200 import _lib
    t = frame.t
    color, radii, depth, flow, acc, idxs = frame._raster(frame.settings(frame.view, frame.proj, frame.campos))
    fn = color.grad_fn
    R = int(fn.num_rendered)
    img = fn.saved_tensors[9]
    cam = frame.sc.cam
    desc = _lib.describe_buffers(frame.P, R, cam.W, cam.H)

    def view(name, dtype):
        b, off, es, cnt = desc[name]
        base = img.data_ptr()
        a = ((base + 255) & ~255) - base
        return img[a + off:a + off + es * cnt].cpu().numpy().view(dtype)

    ranges = view("ranges", np.uint32).reshape(-1, 2).astype(np.int64)
    word = view("tile_batches", np.uint32).astype(np.int64)
    batches, kept = word & 0xFF, word >> 8
    rl = ranges[:, 1] - ranges[:, 0]
    import ctypes
    kb, kw = ctypes.c_int(0), ctypes.c_int(0)
    _lib.load().ex4dgs_forward_geometry(ctypes.byref(kb), ctypes.byref(kw))
    r_eff = int(np.minimum(rl, kb.value * batches).sum())
    n_contrib = view("n_contrib", np.uint32)
    # SURVEY 8d's literal definition (256-splat batches): sum_tiles min(range_len, 256 * ceil(max_pix n_contrib / 256))
    gx, gy = (cam.W + 15) // 16, (cam.H + 15) // 16
    nc = np.zeros((gy * 16, gx * 16), np.int64)
    nc[:cam.H, :cam.W] = n_contrib.reshape(cam.H, cam.W)
    tile_max = nc.reshape(gy, 16, gx, 16).max(axis=(1, 3)).reshape(-1)
    r_eff_256 = int(np.minimum(rl, 256 * ((tile_max + 255) // 256)).sum())
    return dict(R=R, P_vis=int((radii > 0).sum().item()), R_eff=r_eff, R_eff_256=r_eff_256, tiles=int(ranges.shape[0]),
                R_listed=int(rl.sum()), block_keep=float(kept.sum()) / max(1.0, float(kw.value) * r_eff),
                mean_n_contrib=float(n_contrib.mean()))


def cpu_oracle_baseline(sc: synth.Scene):
    """One full fwd+bwd frame of the same workload on the CPU oracle port, all host threads."""
    from oracle import oracle as orc
    from oracle import getters_oracle as GO
    inp = {k: v.numpy() for k, v in GO.flat_inputs(sc).items()}
    go = {k: v.numpy() for k, v in synth.grad_outputs(sc).items()}
    cam = sc.cam
    o = orc.Oracle()
    t0 = time.time()
    o.forward(bg=sc.bg.numpy(), W=cam.W, H=cam.H, means3D=inp["means3D"], dir3D=inp["dir3D"], opacities=inp["opacities"],
              shs=inp["shs"], scales=inp["scales"], rotations=inp["rotations"], viewmatrix=cam.viewmatrix.numpy(),
              projmatrix=cam.projmatrix.numpy(), campos=cam.campos.numpy(), tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
              kernel_size=cam.kernel_size, subpixel_offset=None, min_depth=cam.min_depth, max_depth=cam.max_depth,
              sh_degree=sc.sh_degree)
    o.backward(go["grad_color"], go["grad_depth"], go["grad_flow"], go["grad_acc"])
    dt = time.time() - t0
    return {"value": 1.0 / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "1 full frame (fwd+bwd) of the same workload, oracle/cpu_raster.c with OpenMP, %.1f s" % dt}



def other_configs(mod, dev, K, W, cpu=True):
    """Extra keys on BASELINE.json's other configs (parity-test cases, not the headline): config 1 on the CPU oracle port
    (the config the reference's CPU plumbing would run), config 2 = 500 k static Gaussians forward-only, and the frame
    size of config 5 (Technicolor 2048x1088, configs/techni/Painter.json:2) with the C3 Gaussian counts, fwd+bwd."""
    out = {}

    def run(sc, fwd_only):
        fr = Frame(mod, sc, dev, seed_off=0, impl="ours")
        fn = fr.step_forward if fwd_only else fr.step_device
        for _ in range(max(W, 3)):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        st = frame_stats(fr)
        del fr
        torch.cuda.empty_cache()
        return {"value": 1000.0 / ms, "unit": "frames/s", "ms_per_step": ms, "steps": K, "P": sc.P, "P_vis": st["P_vis"], "R": st["R"]}

    sc2 = synth.make_config("C2")
    out["config2_fwd_only_500k_1352x1014"] = run(sc2, True)
    sc5 = synth.make_scene(1_500_000, 500_000, 2048, 1088)
    out["config5_frame_2048x1088_fwd_bwd"] = run(sc5, False)
    if cpu:
        sc1 = synth.make_config("C1")
        b = cpu_oracle_baseline(sc1)
        b["sample"] = "config 1 (10 k static Gaussians, 400x400), " + b["sample"]
        out["config1_cpu_port"] = b
    return out



def densify_leg(sc: synth.Scene, dev, impl: str, reps: int = 3):
    """The tensor surgery of densification at this workload (row N4): the reference's UNMODIFIED CGaussianModel - its own
    constructor, training_setup (torch.optim.RAdam over the 15 named groups, arguments/__init__.py:93-111 learning rates)
    and one optimizer step so that both moments exist - runs `prune_points` on a random 10 % of the Gaussians and then
    `densification_postfix` with 5 % cloned rows (scene/c_gaussian_model.py:715-844); reference arm: the class's own
    `_prune_optimizer` / `cat_tensors_to_optimizer` (one x[mask] or torch.cat per tensor and per moment), our arm:
    ex4dgs_b200.densify bound onto the same object (one ex4dgs_gather_rows launch per call).  Host wall clock incl. the
    synchronisation the masks need, best of `reps`."""
    from types import SimpleNamespace
    cls = load_reference_model_class()
    if cls is None:
        return {"unavailable": "oracle/_ref/callers/scene/c_gaussian_model.py is missing"}
    args = SimpleNamespace(
        percent_dense=0.01, position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01,
        position_lr_max_steps=30000, dynamic_position_lr_init=0.00016, dynamic_position_lr_final=0.000016,
        dynamic_position_lr_delay_mult=0.01, dynamic_position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.05,
        scaling_lr=0.005, rotation_lr=0.00001, disp_lr=0.0001, feature_motion_lr=0.0025, rotation_motion_lr=0.001,
        opacity_motion_lr=0.05, opacity_motion_center_lr=0.001, opacity_motion_var_lr=0.0005)
    m = reference_model(sc, dev, cls)
    m.spatial_lr_scale = 1.0
    m.keyframe_num = int(m._xyz_motion.shape[1])
    prev = torch.cuda.current_device()
    torch.cuda.set_device(dev)          # the class allocates with device="cuda"
    try:
        m.training_setup(args)
        m.max_radii2D = torch.zeros(m._xyz.shape[0], device=dev)
        m.min_radii2D = torch.ones(m._xyz.shape[0], device=dev) * 1000
        m.motion_max_radii2D = torch.zeros(m._xyz_motion.shape[0], device=dev)
        m.motion_min_radii2D = torch.ones(m._xyz_motion.shape[0], device=dev) * 1000
        params = [g["params"][0] for g in m.optimizer.param_groups]
        for p in params:
            p.grad = torch.full_like(p, 1e-4)
        m.optimizer.step()
        m.optimizer.zero_grad(set_to_none=True)
        del params
        if impl == "ours":
            from ex4dgs_b200 import densify
            densify.install(m)
        static_names = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "_xyz_disp")
        dynamic_names = ("_xyz_motion", "_features_dc_motion", "_features_rest_motion", "_scaling_motion", "_opacity_motion",
                         "_opacity_duration_center", "_opacity_duration_var", "_rotation_motion")
        gen = torch.Generator(device=dev).manual_seed(5)
        t_prune, t_cat = [], []
        for _ in range(reps):
            ms = torch.rand(m._xyz.shape[0], generator=gen, device=dev) < 0.1
            md = torch.rand(m._xyz_motion.shape[0], generator=gen, device=dev) < 0.1
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            m.prune_points(ms, md)
            torch.cuda.synchronize(dev)
            t_prune.append((time.perf_counter() - t0) * 1e3)
            sel_s = torch.nonzero(torch.rand(m._xyz.shape[0], generator=gen, device=dev) < 0.05).squeeze(1)
            sel_d = torch.nonzero(torch.rand(m._xyz_motion.shape[0], generator=gen, device=dev) < 0.05).squeeze(1)
            with torch.no_grad():
                new = [getattr(m, n)[sel_s] for n in static_names] + [getattr(m, n)[sel_d] for n in dynamic_names]
                # the two statistics densification_postfix does not reset are extended by its callers (c_gaussian_model.py:985-988,1008-1011)
                for n, sel in (("xyz_error_min", sel_s), ("xyz_error_min_timestamp", sel_s),
                               ("motion_xyz_error_min", sel_d), ("motion_xyz_error_min_timestamp", sel_d)):
                    setattr(m, n, torch.cat([getattr(m, n), getattr(m, n)[sel]]))
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            m.densification_postfix(*new)
            torch.cuda.synchronize(dev)
            t_cat.append((time.perf_counter() - t0) * 1e3)
            del new
        n_par = sum(int(g["params"][0].numel()) for g in m.optimizer.param_groups)
    finally:
        torch.cuda.set_device(prev)
    del m
    torch.cuda.empty_cache()
    return {"prune_points_ms": min(t_prune), "densification_postfix_ms": min(t_cat), "parameters": n_par, "reps": reps,
            "what": "CGaussianModel.prune_points (10 % removed) and densification_postfix (5 % appended) on the reference class: "
                    + ("ex4dgs_b200.densify bound on (one gather launch per call)" if impl == "ours" else "its own per-tensor x[mask] / torch.cat")}


def config4_sweep(sc: synth.Scene, dev, rank: int, ws: int, frames: int = 300):
    """BASELINE.json config 4 in synthetic form, driver-visible: the render.py sweep (render.py:64-96) - `frames` frames,
    one timestamp each, forward only - sharded frame i -> rank i mod N.  The model goes through the reference's on-disk
    format on the way: rank 0 writes the scene as point_cloud.ply + dynamic_point_cloud.ply (ex4dgs_b200/model_io.py,
    byte-compatible with CGaussianModel.save_ply), every rank loads the two files, wraps the loaded arrays in FusedGetters
    and renders its share through GaussianRasterizer (fused front-end kernel + rasterizer forward per frame)."""
    import shutil
    import tempfile
    from ex4dgs_b200 import model_io, parallel
    from ex4dgs_b200.frontend import FusedGetters
    import ex4dgs_b200 as m
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    d = os.path.join(base, "ex4dgs_b200_bench_model_%s" % os.environ.get("MASTER_PORT", "single"))
    path = os.path.join(d, "point_cloud.ply")
    t_io = time.time()
    if rank == 0:
        os.makedirs(d, exist_ok=True)
        ga = model_io.GaussianArrays(
            _xyz=sc.xyz, _features_dc=sc.features[:, :1], _features_rest=sc.features[:, 1:], _opacity=sc.opacity,
            _scaling=sc.scaling, _rotation=sc.rotation, _xyz_disp=sc.xyz_disp, _xyz_motion=sc.xyz_motion,
            _features_dc_motion=sc.features_motion[:, :1], _features_rest_motion=sc.features_motion[:, 1:],
            _scaling_motion=sc.scaling_motion, _opacity_motion=sc.opacity_motion,
            _opacity_duration_center=sc.opacity_center[:, :, None], _opacity_duration_var=sc.opacity_var[:, :, None],
            _rotation_motion=sc.rotation_motion, duration=sc.duration, interval=sc.interval, time_pad=sc.time_pad,
            time_shift=sc.time_shift, var_pad=sc.var_pad, kernel_size=sc.cam.kernel_size)
        model_io.save_model(ga, path)
    if ws > 1:
        torch.distributed.barrier()
    ga = model_io.load_model(path, 3, duration=sc.duration, interval=sc.interval, time_pad=sc.time_pad, var_pad=sc.var_pad,
                             kernel_size=sc.cam.kernel_size, device=dev, check_keyframes=False)
    nbytes = os.path.getsize(path) + os.path.getsize(path.replace("point_cloud.ply", "dynamic_point_cloud.ply"))
    t_io = time.time() - t_io
    if ws > 1:
        torch.distributed.barrier()
    if rank == 0:
        shutil.rmtree(d, ignore_errors=True)
    cam = sc.cam
    pc = FusedGetters(ga)
    rs = m.GaussianRasterizationSettings(
        image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, kernel_size=cam.kernel_size,
        subpixel_offset=torch.zeros(cam.H, cam.W, 2, device=dev), bg=sc.bg.to(dev), scale_modifier=1.0,
        viewmatrix=cam.viewmatrix.to(dev), projmatrix=cam.projmatrix.to(dev), sh_degree=sc.sh_degree,
        campos=cam.campos.to(dev), prefiltered=False, min_depth=cam.min_depth, max_depth=cam.max_depth, debug=False)
    rast = m.GaussianRasterizer(rs)
    zeros3 = torch.zeros(ga.num_static + ga.num_dynamic, 3, device=dev)
    timestamps = [float(i % 300) for i in range(frames)]
    mine = parallel.shard_frames(len(timestamps), rank, ws)

    def frame(t):
        with torch.no_grad():
            # the order render() calls them in (gaussian_renderer/__init__.py:62-95): get_features() ends the frame
            return rast(means3D=pc.get_xyz_at_t(t), means2D=zeros3, dir3D=zeros3, opacities=pc.get_opacity_at_t(t),
                        scales=pc.get_scaling(), rotations=pc.get_rotation_at_t(t), shs=pc.get_features())[0]

    for t in timestamps[:3]:
        frame(t)
    torch.cuda.synchronize()
    if ws > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    checksum = torch.zeros((), device=dev, dtype=torch.float64)
    for i in mine:
        checksum += frame(timestamps[i]).double().sum()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if ws > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(checksum)
    return {"value": len(timestamps) / (ms.item() / 1e3), "unit": "frames/s", "frames": len(timestamps), "ms_total": ms.item(),
            "n_gpus": ws, "model_files_bytes": int(nbytes), "model_io_s": round(t_io, 2), "checksum": checksum.item(),
            "what": "render.py-style sweep: 300 timestamps, forward only, frame i -> rank i mod N; model written to and loaded "
                    "from the reference's two-file PLY format; per frame: fused front-end kernel (SegmentedSH, no torch.cat) "
                    "+ rasterizer forward"}



def config5_train_step(mod, dev, rank, ws, group, K, W):
    """BASELINE.json config 5 in synthetic form, driver-visible at every N: the training step of train.py:139-172 (H2D camera
    + ground truth, render, 0.8 L1 + 0.2 (1 - SSIM) + the l1_accum hook tensor, backward, loss all-reduce + D2H) at the
    Technicolor frame size 2048x1088 (configs/techni/Painter.json:2), ONE camera per rank per step - 8 cameras per step,
    data-parallel, loss all-reduce only, when launched on 8 GPUs.  Gaussian counts of C3 (BASELINE.json fixes only those)."""
    sc5 = synth.make_scene(1_500_000, 500_000, 2048, 1088)
    fr = Frame(mod, sc5, dev, seed_off=rank, impl="ours")
    step = make_train_step(fr, None)
    for _ in range(max(W, 3)):
        step(group)
    torch.cuda.synchronize()
    if ws > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step(group)
    e1.record()
    torch.cuda.synchronize()
    if ws > 1:
        torch.distributed.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if ws > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    out = {"value": ws * K / (ms.item() / 1e3), "unit": "cameras/s", "ms_per_step": ms.item() / K, "steps": K, "n_gpus": ws,
           "cameras_per_step": ws, "image": "2048x1088", "P": sc5.P,
           "what": "train.py step (render + photometric loss + backward + loss all-reduce), one camera per rank per step"}
    del step, fr
    torch.cuda.empty_cache()
    return out


_JSON_FD = None


def _protect_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints
    "NCCL version ..." to fd 1 under NCCL_DEBUG=VERSION/INFO): keep a private copy of the real stdout for
    the JSON line and point fd 1 at stderr for everything else."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cpu"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true")
    ap.add_argument("--fwd-only", action="store_true",
                    help="forward-only render under no_grad (BASELINE.json config 2); changes the metric name")
    ap.add_argument("--profile-in-timed", type=int, default=1,
                    help="record per-stage CUDA events inside the timed region (1) or in a separate pass (0)")
    ap.add_argument("--no-train-iter", action="store_true", help="skip the full-training-iteration leg")
    ap.add_argument("--only-train-iter", action="store_true", help="profiling aid: run only the train_iter leg")
    ap.add_argument("--bookkeeping", type=int, default=1,
                    help="--only-train-iter: include the per-iteration regularisers / statistics / NaN guards (train_iter_full)")
    ap.add_argument("--dp-grads", action="store_true",
                    help="train_iter leg under torchrun: SUM all-reduce of all gradients before the optimizer step")
    ap.add_argument("--l2-flush", action="store_true",
                    help="SURVEY 8d: write a 256 MB buffer (> the 126 MB L2) between the timed iterations of the device-timed leg; "
                         "each iteration is then timed with its own pair of CUDA events and the flush is excluded")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip the extra keys measured on BASELINE.json's other configs (1: CPU, 2: forward-only 500k, 5: 2048x1088)")
    ap.add_argument("--value-only", action="store_true", help="tuning aid (tools/tune.py): only the device-timed leg + stage timers")
    ap.add_argument("--tile-cull", type=int, default=int(os.environ.get("EX4DGS_TILE_CULL", "1")))
    args = ap.parse_args()
    rank, local_rank, ws = dist_env()
    K, W = args.steps, max(args.warmup, 3)
    if ws > 1:
        # one slice of the host cores per rank (all 8 GPUs of these boxes hang off one NUMA node): the per-frame Python
        # work of 8 ranks otherwise migrates between cores and shows up as e2e jitter
        try:
            n = os.cpu_count() or 1
            per = max(1, n // ws)
            os.sched_setaffinity(0, set(range(local_rank * per, min(n, (local_rank + 1) * per))))
            torch.set_num_threads(max(1, min(4, per)))
        except Exception:
            pass
    sc = synth.make_config(args.workload)
    cam = sc.cam
    cfg = {"workload": "%s: %d static + %d dynamic Gaussians (K=36 keyframes, pre-interpolated at t=137), %dx%d, SH degree 3, "
                       "fwd+bwd at the GaussianRasterizer boundary" % (args.workload, sc.xyz.shape[0], sc.xyz_motion.shape[0], cam.W, cam.H),
           "parallelism": "frame-parallel x%d (Gaussians replicated, one frame per GPU per step)" % ws,
           "l2": ("256 MB written between timed iterations (--l2-flush), each iteration timed by its own event pair"
                  if args.l2_flush else "inputs (496 MB at C3) larger than the 126 MB L2; no explicit flush (--l2-flush adds one)")}

    # ---------------- reference arm on the CPU (fallback when oracle/_ref is absent) ----------------
    ref_mod = None
    if args.impl != "ours":
        ref_mod = load_reference() if args.impl == "reference" and torch.cuda.is_available() else None
        if ref_mod is None:
            if rank != 0:
                return
            cb = cpu_oracle_baseline(sc)
            line = {"metric": METRIC, "value": cb["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
                    "ms_per_step": 1000.0 / cb["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "f32", "data": "synthetic", "config": cfg, "impl": "reference", "cpu_baseline": cb,
                    "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0,
                    "note": "reference CUDA extension not available in oracle/_ref; CPU oracle port timed instead"}
            emit(line)
            return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if ws > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
        group = torch.distributed.group.WORLD

    if args.impl == "ours":
        import ex4dgs_b200 as mod
        from ex4dgs_b200 import _lib
        mod.set_default_flags(bool(args.tile_cull))
        lib = _lib.load()
    else:
        mod, lib = ref_mod, None

    frame = Frame(mod, sc, dev, seed_off=rank, impl=args.impl if args.impl == "ours" else "reference")
    if args.fwd_only:
        frame.step_device = frame.step_forward
        frame.step_e2e = lambda group=None: frame.step_forward()

    def barrier():
        torch.cuda.synchronize()
        if ws > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(256 * 1024 * 1024 // 4, device=dev) if args.l2_flush else None

    def timed(step_fn, steps, flush=False):
        barrier()
        t0 = time.time()
        if flush and flush_buf is not None:
            evs = []
            for _ in range(steps):
                flush_buf.fill_(1.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                step_fn()
                b.record()
                evs.append((a, b))
            barrier()
            t1 = time.time()
            ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev)
            if ws > 1:
                torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
            return float(ms.item()), t0, t1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_fn()
        e1.record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if ws > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    if args.only_train_iter:
        # profiling aid (ncu launch lists): nothing but the full-training-iteration leg
        rl = None
        if args.impl != "ours":
            rl = load_reference_loss()
        model_step, n_param = make_model_step(frame, "ours" if args.impl == "ours" else "reference", rl, args.dp_grads,
                                              bookkeeping=bool(args.bookkeeping),
                                              ref_model_cls=load_reference_model_class() if args.impl != "ours" else None)
        for _ in range(W):
            model_step(group)
        ms_iter, _, _ = timed(lambda: model_step(group), K)
        if rank == 0:
            emit({"train_iter_ms": ms_iter / K, "steps": K, "parameters": n_param, "impl": args.impl,
                  "bookkeeping": bool(args.bookkeeping)})
        if ws > 1:
            torch.distributed.destroy_process_group()
        return

    # warm-up: W steps of each leg, then keep going until the clocks have ramped (~1 s of work)
    for _ in range(W):
        frame.step_device()
    t_spin = time.time()
    while time.time() - t_spin < 1.0:
        frame.step_device()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if (rank == 0 and not args.no_clocks) else None
    time.sleep(0.5 if sampler else 0.0)
    for _ in range(3):
        frame.step_device()

    def read_stages():
        lib.ex4dgs_profile_enable(0)
        arr = (C.c_double * 6)()
        nf, nb = C.c_int(0), C.c_int(0)
        lib.ex4dgs_profile_read(arr, C.byref(nf), C.byref(nb))
        return [arr[i] / max(1, (nf.value if i < 4 else nb.value)) for i in range(6)]

    launches0 = lib.ex4dgs_launch_count() if lib else 0
    if lib and args.profile_in_timed:
        lib.ex4dgs_profile_enable(1)
    ms_total, t0, t1 = timed(frame.step_device, K, flush=args.l2_flush)
    launches = (lib.ex4dgs_launch_count() - launches0) if lib else 0
    stage_ms = None
    if lib:
        if not args.profile_in_timed:
            lib.ex4dgs_profile_enable(1)
            timed(frame.step_device, min(K, 20))
        stage_ms = read_stages()
    clocks = sampler.stop(t0, t1) if sampler else None
    if args.value_only:
        if rank == 0:
            names = ["preprocess_fwd", "depth_sort_scan", "sync_duplicate_tilesort_ranges", "render_fwd", "render_bwd", "preprocess_bwd"]
            emit({"value": ws * K / (ms_total / 1000.0), "ms_per_step": ms_total / K, "steps": K,
                  "stage_ms": dict(zip(names, stage_ms)) if stage_ms else None, "clocks": clocks, "tuning_run": True})
        if ws > 1:
            torch.distributed.destroy_process_group()
        return

    # end-to-end leg
    for _ in range(W):
        frame.step_e2e(group)
    ms_e2e, _, _ = timed(lambda: frame.step_e2e(group), K)

    # the frame through the reference's unchanged gaussian_renderer.render()
    render_api = None
    if not args.fwd_only:
        rstep = make_render_api_step(frame, "ours" if args.impl == "ours" else "reference", mod)
        if rstep is not None:
            Kr = max(20, K // 3)
            for _ in range(W):
                rstep()
            ms_r, _, _ = timed(rstep, Kr)
            render_api = {"value": ws * Kr / (ms_r / 1000.0), "unit": "frames/s", "ms_per_step": ms_r / Kr, "steps": Kr,
                          "what": "fwd+bwd through the UNMODIFIED gaussian_renderer.render() of the reference (per-frame "
                                  "getters, its torch.cuda.synchronize()), gradients down to the model parameters; "
                                  + ("FusedGetters + this repository's drop-in package" if args.impl == "ours" else
                                     "the reference's CGaussianModel getters + its own extension")}
            del rstep
            torch.cuda.empty_cache()

    # training-iteration leg (row N2): the same step with the reference's loss block, train.py:144-151
    train = None
    ref_loss = None
    if not args.fwd_only:
        if args.impl != "ours":
            ref_loss = load_reference_loss()
            if ref_loss is None:                       # restated in float32 torch ops by the oracle
                from oracle import loss_oracle
                class _L:                              # noqa: E306
                    l1_loss = staticmethod(lambda a, b: (a - b).abs().mean())
                    @staticmethod
                    def ssim(a, b, reduce=True):
                        m = loss_oracle.ssim_map(a, b, torch.float32)
                        return m.mean() if reduce else m
                ref_loss = _L
        train_step = make_train_step(frame, ref_loss)
        for _ in range(W):
            train_step(group)
        ms_train, _, _ = timed(lambda: train_step(group), K)
        train = {"value": ws * K / (ms_train / 1000.0), "unit": "frames/s", "ms_per_step": ms_train / K,
                 "what": "training iteration around the rasterizer (train.py:139-172): H2D camera + ground truth, render, "
                         "loss = 0.8 L1 + 0.2 (1 - SSIM) + l1_accum hook tensor, backward, loss D2H; "
                         + ("fused loss kernels (ex4dgs_b200/loss.py)" if args.impl == "ours" else
                            "the reference's utils/loss_utils.py (torch conv2d)")}

    # full training iteration on the model's native parameters (rows N1 + N2 + N4 around the path)
    train_iter = train_iter_full = None
    if not args.fwd_only and not args.no_train_iter:
        Kt = max(20, K // 3)
        ref_cls = load_reference_model_class() if args.impl != "ours" else None
        if args.impl == "ours" or ref_cls is not None:
            model_step, n_param = make_model_step(frame, "ours" if args.impl == "ours" else "reference", ref_loss, args.dp_grads,
                                                  bookkeeping=True, ref_model_cls=ref_cls)
            for _ in range(W):
                model_step(group)
            ms_full, _, _ = timed(lambda: model_step(group), Kt)
            train_iter_full = {"value": ws * Kt / (ms_full / 1000.0), "unit": "iterations/s", "ms_per_step": ms_full / Kt, "steps": Kt,
                               "what": "train_iter + everything else train.py does every iteration outside densify_and_prune: "
                                       "static / motion regularisation terms (:156-162), mark_prune_stats, max_radii2D, "
                                       "add_densification_stats, add_l1_ssim_stats (:196-212), nan_to_num of one gradient (:246-248), "
                                       "prune_nan_points' NaN test (:253); "
                                       + ("fused statistics kernel + regulariser kernel + guards inside the RAdam kernel"
                                          if args.impl == "ours" else
                                          "the reference's own CGaussianModel methods and train.py expressions")}
            del model_step
            torch.cuda.empty_cache()
        model_step, n_param = make_model_step(frame, "ours" if args.impl == "ours" else "reference", ref_loss, args.dp_grads,
                                              bookkeeping=False)
        for _ in range(W):
            model_step(group)
        ms_iter, _, _ = timed(lambda: model_step(group), Kt)
        train_iter = {"value": ws * Kt / (ms_iter / 1000.0), "unit": "iterations/s", "ms_per_step": ms_iter / Kt, "steps": Kt,
                      "parameters": n_param, "lr_scale": LR_SCALE, "gradient_allreduce": bool(args.dp_grads and ws > 1), "gradient_allreduce_overlapped": bool(DP_OVERLAP),
                      "what": "train.py iteration without densification on the 15 native parameter tensors: per-frame getters, "
                              "get_features, render, loss block, backward to the parameters, RAdam step, zero_grad; "
                              + ("fused front-end + segmented SH input (no torch.cat) + fused loss + FusedRAdam (one kernel each)"
                                 if args.impl == "ours" else
                                 "PyTorch getters + get_features torch.cat + utils/loss_utils.py + torch.optim.RAdam (foreach)")}

    sweep4 = train5 = None
    if args.impl == "ours" and not args.fwd_only and not args.no_extra_configs and args.workload == "C3":
        sweep4 = config4_sweep(sc, dev, rank, ws)
        train5 = config5_train_step(mod, dev, rank, ws, group, min(K, 60), W)

    value = ws * K / (ms_total / 1000.0)
    e2e_value = ws * K / (ms_e2e / 1000.0)
    if rank != 0:
        if ws > 1:
            torch.distributed.destroy_process_group()
        return

    h2d = frame.h_gt.numel() * 4 + frame.h_cam.numel() * 4
    line = {"metric": METRIC if not args.fwd_only else "fwd-only frames/sec (config 2 style)", "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / K,
                    "what": "per frame: H2D camera (35 floats) + ground-truth image from pinned memory, forward, L1 loss "
                            + ("(ex4dgs_b200.loss.l1_loss)" if args.impl == "ours" else "(utils/loss_utils.py l1_loss: torch abs/mean)")
                            + ", backward, loss (all-reduced over ranks) D2H; Gaussian parameters resident"},
            "clocks": clocks}
    if render_api is not None:
        line["render_api"] = render_api
    if sweep4 is not None:
        line["config4_render_sweep"] = sweep4
    if train5 is not None:
        line["config5_train_step_2048x1088"] = train5
    if train is not None:
        line["train_step"] = train
    if train_iter is not None:
        line["train_iter"] = train_iter
    if train_iter_full is not None:
        line["train_iter_full"] = train_iter_full
    if args.impl == "ours":
        st = frame_stats(frame)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
        alg_bytes = 56 * st["R_eff"] + 52 * cam.W * cam.H + 8 * st["tiles"]
        t_render = stage_ms[3] / 1000.0
        achieved = alg_bytes / t_render / 1e9
        line["gpu_launches"] = int(launches)
        line["scene_stats"] = {"P": frame.P, "P_vis": st["P_vis"], "R": st["R"], "R_eff": st["R_eff"], "R_eff_256": st["R_eff_256"],
                               "R_listed": st["R_listed"], "block_keep": st["block_keep"],
                               "tile_cull": int(args.tile_cull), "mean_n_contrib": st["mean_n_contrib"]}
        traffic, traffic_src = None, "no ncu capture committed"
        tj_all = None
        try:
            tj_all = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj_all.get("kernel_source_sha16") != kernel_source_hash() or args.workload != "C3" or not args.tile_cull:
                traffic_src = ("profiles/traffic.json was captured on other kernel sources / another workload (%s): not used"
                               % tj_all.get("kernel_source_sha16"))
                tj_all = None
            else:
                traffic, traffic_src = tj_all["render_fwd_kernel"]["dram_bytes_per_launch"], tj_all["render_fwd_kernel"]["source"]
        except Exception:
            tj_all = None
        line["roofline"] = {"bound": "hbm", "kernel": "render_fwd_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                            "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": stage_ms[3],
                            "R_eff": "instances in the 128-splat batches the tiles actually fetch (scene_stats.R_eff); with SURVEY 8d's "
                                     "literal 256-splat granularity (scene_stats.R_eff_256) the fraction is frac_at_256",
                            "frac_at_256": (56 * st["R_eff_256"] + 52 * cam.W * cam.H + 8 * st["tiles"]) / t_render / 1e9 / peak,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                            "note": "the kernel is bound by instruction issue / the FP32 and ALU pipes, not by HBM (ncu: 5 % DRAM): see issue_slots and profiles/SUMMARY.md"}
        # what actually bounds the two compositing kernels: warp instructions issued (smsp__inst_executed.sum of the committed
        # ncu capture - a property of the workload, identical in every launch) over the live kernel time, against the SMs'
        # issue rate (4 schedulers x 1 warp instruction per clock x SM count x the SM clock measured under load)
        try:
            tj = tj_all
            if tj is None:
                raise KeyError("no valid capture")
            props = torch.cuda.get_device_properties(dev)
            mhz = (clocks or {}).get("sm_mhz") or 1965.0
            peak_issue = props.multi_processor_count * 4 * mhz * 1e6
            line["issue_slots"] = {
                k: {"warp_inst_per_launch": tj[k + "_kernel"]["inst_executed_per_launch"], "kernel_ms": ms,
                    "achieved_Ginst_s": tj[k + "_kernel"]["inst_executed_per_launch"] / (ms * 1e-3) / 1e9,
                    "peak_Ginst_s": peak_issue / 1e9,
                    "frac": tj[k + "_kernel"]["inst_executed_per_launch"] / (ms * 1e-3) / peak_issue}
                for k, ms in (("render_fwd", stage_ms[3]), ("render_bwd", stage_ms[4])) if ms > 0 and args.workload == "C3" and args.tile_cull}
        except Exception:
            pass
        names = ["preprocess_fwd", "depth_sort_scan", "sync_duplicate_tilesort_ranges", "render_fwd", "render_bwd", "preprocess_bwd"]
        line["stage_ms"] = dict(zip(names, stage_ms))
        # algorithmic bytes per stage (with this design's record sizes) against the same HBM peak
        Pn, Pv, Rn, Re, px = frame.P, st["P_vis"], st["R"], st["R_eff"], cam.W * cam.H
        stage_bytes = [20 * Pn + 291 * Pv,                       # means + always-written words; visible: inputs + 64 B record
                       (2 * 8 * 4 + 8) * Pn,                     # 4 radix passes over (u32 key, u32 value) + the scan
                       8 * Pn + 12 * Pv + 6 * Rn + (2 * 6 * 2 + 2) * Rn + 2 * Rn + 8 * st["tiles"],   # duplicate, 2 passes of (u16, u32), ranges
                       alg_bytes,
                       44 * Re + 56 * px + 56 * Pv,              # staged records, per-pixel inputs, gradient read-modify-write
                       579 * Pv + 156 * Pn]                      # inputs + accumulator of visible, dense zero-filled outputs
        line["stage_roofline"] = {n: {"algorithmic_bytes": int(b), "GBps": b / (ms * 1e-3) / 1e9, "frac": b / (ms * 1e-3) / 1e9 / peak}
                                  for n, b, ms in zip(names, stage_bytes, stage_ms) if ms > 0}
        if not args.no_cpu_baseline and ws == 1:
            line["cpu_baseline"] = cpu_oracle_baseline(sc)
        if not args.no_extra_configs and ws == 1 and not args.fwd_only and args.workload == "C3":
            line["other_configs"] = other_configs(mod, dev, min(K, 60), W, cpu=not args.no_cpu_baseline)
    else:
        line["impl"] = "reference"
        line["gpu_launches"] = 0
        line["cpu_baseline"] = {"value": value, "unit": "frames/s", "cores": 0, "kind": "reference",
                                "sample": "the unmodified reference CUDA extension (oracle/_ref) on the same GPU - the reference "
                                          "path has no CPU implementation (rasterize_points.cu:80)"}
    if ws == 1 and args.workload == "C3" and not (args.no_extra_configs or args.value_only or args.fwd_only or args.only_train_iter):
        try:
            line["densify_and_prune"] = densify_leg(sc, dev, "ours" if args.impl == "ours" else "reference")
        except Exception as e:      # noqa: BLE001 - an extra leg must never cost the headline line
            line["densify_and_prune"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    emit(line)
    if ws > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
