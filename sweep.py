#!/usr/bin/env python
"""Config 4 of BASELINE.json in synthetic form: a 300-frame render sweep (forward only, one
timestamp per frame) sharded round-robin over the GPUs of one node, the way the reference's
render.py loop (render.py:64-88) would be sharded.  Per frame: the fused front-end (row N1: keyframe
interpolation of the dynamic Gaussians at that timestamp, one kernel) + the rasterizer forward.
Gaussians are replicated, no collective on the render path; the per-rank times are max-reduced.

    python sweep.py [--frames 300] [--workload C3]
    python sweep.py --model-path <trained model dir> [--iteration -1] --duration 300 --interval 10 --time-pad 2
        (row N3: the reference's point_cloud.ply + dynamic_point_cloud.ply loaded by ex4dgs_b200/model_io.py;
         the camera stays the synthetic one - dataset readers are out of scope)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 sweep.py
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from ex4dgs_b200 import parallel, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--fused", type=int, default=1, help="1: fused front-end kernel, 0: comparison arm - the PyTorch restatement of the reference getters (oracle/getters_oracle.py)")
    ap.add_argument("--model-path", default=None, help="trained reference model directory (point_cloud/iteration_*/...)")
    ap.add_argument("--iteration", type=int, default=-1)
    ap.add_argument("--duration", type=float, default=300.0)
    ap.add_argument("--interval", type=float, default=10.0)
    ap.add_argument("--time-pad", type=float, default=2.0)
    args = ap.parse_args()
    rank, local_rank, ws = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if ws > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import ex4dgs_b200 as m
    from ex4dgs_b200.frontend import interpolate_gaussians

    sc = synth.make_config(args.workload if args.model_path is None else "C1")
    cam = sc.cam
    if args.model_path is not None:
        # swap the synthetic Gaussians for a trained model in the reference's on-disk format
        import copy
        from ex4dgs_b200 import model_io
        ga, it = model_io.load_iteration(args.model_path, args.iteration, duration=args.duration, interval=args.interval,
                                         time_pad=args.time_pad)
        sc = copy.copy(sc)
        cam = copy.copy(cam)
        cam.W, cam.H = 1352, 1014
        cam.tanfovx, cam.tanfovy = cam.W / (2 * 1462.0), cam.H / (2 * 1462.0)
        for dst, src in (("xyz", "_xyz"), ("xyz_disp", "_xyz_disp"), ("rotation", "_rotation"), ("scaling", "_scaling"),
                         ("opacity", "_opacity"), ("xyz_motion", "_xyz_motion"), ("rotation_motion", "_rotation_motion"),
                         ("scaling_motion", "_scaling_motion"), ("opacity_motion", "_opacity_motion"),
                         ("opacity_center", "_opacity_duration_center"), ("opacity_var", "_opacity_duration_var")):
            setattr(sc, dst, getattr(ga, src))
        sc.features = torch.cat((ga._features_dc, ga._features_rest), dim=1)
        sc.features_motion = torch.cat((ga._features_dc_motion, ga._features_rest_motion), dim=1)
        sc.duration, sc.interval, sc.time_shift, sc.var_pad = ga.duration, ga.interval, ga.time_shift, ga.var_pad
        sc.cam = cam
    names = ["xyz", "xyz_disp", "rotation", "scaling", "opacity", "xyz_motion", "rotation_motion", "scaling_motion",
             "opacity_motion", "opacity_center", "opacity_var"]
    T = {n: getattr(sc, n).to(dev) for n in names}
    shs = torch.cat([sc.features, sc.features_motion]).to(dev)          # timestamp-independent: concatenated once
    P = shs.shape[0]
    zeros3 = torch.zeros(P, 3, device=dev)
    rs = m.GaussianRasterizationSettings(
        image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, kernel_size=cam.kernel_size,
        subpixel_offset=torch.zeros(cam.H, cam.W, 2, device=dev), bg=sc.bg.to(dev), scale_modifier=1.0,
        viewmatrix=cam.viewmatrix.to(dev), projmatrix=cam.projmatrix.to(dev), sh_degree=sc.sh_degree,
        campos=cam.campos.to(dev), prefiltered=False, min_depth=cam.min_depth, max_depth=cam.max_depth, debug=False)
    rast = m.GaussianRasterizer(rs)
    import copy
    sc_dev = copy.copy(sc)
    for n in names:
        setattr(sc_dev, n, T[n])
    sc_dev.features, sc_dev.features_motion = shs[:sc.xyz.shape[0]], shs[sc.xyz.shape[0]:]
    timestamps = [float(i % 300) for i in range(args.frames)]
    mine = parallel.shard_frames(len(timestamps), rank, ws)

    def frame(t):
        with torch.no_grad():
            if args.fused:
                means, rots, scales, opac = interpolate_gaussians(*[T[n] for n in names], t=t, duration=sc.duration,
                                                                  interval=sc.interval, time_shift=sc.time_shift,
                                                                  var_min=sc.var_pad / sc.interval)
            else:
                from oracle import getters_oracle as GO     # comparison arm only (--fused 0)
                fi = GO.flat_inputs(sc_dev, t)      # PyTorch getters (restatement of c_gaussian_model.py) on the GPU tensors
                means, rots, scales, opac = fi["means3D"], fi["rotations"], fi["scales"], fi["opacities"]
            return rast(means3D=means, means2D=zeros3, dir3D=zeros3, opacities=opac, shs=shs, scales=scales, rotations=rots)[0]

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
    if rank == 0:
        print(json.dumps({"metric": "render-sweep frames/s (config 4 style, forward only)", "value": len(timestamps) / (ms.item() / 1e3),
                          "unit": "frames/s", "n_gpus": ws, "frames": len(timestamps), "ms_total": ms.item(),
                          "front_end": "fused kernel" if args.fused else "PyTorch getters", "checksum": checksum.item(),
                          "config": {"workload": args.workload, "sharding": "frame i -> rank i mod N, Gaussians replicated"}}))
    if ws > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
