"""Analysis tool (CPU only; not product, not a test): how well do different work mappings of the forward compositing
kernel fill a warp on the benchmark scene?  Replays one frame of workload C3 from the CPU oracle's tile lists
(oracle/cpu_raster.c) and counts warp-level executions of the kernel's code sections (tools/simt_sim.c).

    python tools/simt_sim.py [--workload C3]
"""
import argparse
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ex4dgs_b200 import synth  # noqa: E402
from oracle import getters_oracle as GO  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C3")
    args = ap.parse_args()
    so = os.path.join(ROOT, "tools", "libsimt_sim.so")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-ffp-contract=off", "-fopenmp", "-I", os.path.join(ROOT, "oracle"),
                           "-o", so, os.path.join(ROOT, "tools", "simt_sim.c"), "-lm"])
    lib = C.CDLL(so)
    lib.simulate.argtypes = [C.c_int] * 6 + [C.c_void_p] * 5
    lib.simulate_cursor.argtypes = [C.c_int] * 5 + [C.c_void_p] * 5
    lib.simulate_queue.argtypes = [C.c_int] * 6 + [C.c_void_p] * 5
    sc = synth.make_config(args.workload)
    inp = {k: v.numpy() for k, v in GO.flat_inputs(sc).items()}
    cam = sc.cam
    o = orc.Oracle()
    t0 = time.time()
    o.forward(bg=sc.bg.numpy(), W=cam.W, H=cam.H, means3D=inp["means3D"], dir3D=inp["dir3D"], opacities=inp["opacities"],
              shs=inp["shs"], scales=inp["scales"], rotations=inp["rotations"], viewmatrix=cam.viewmatrix.numpy(),
              projmatrix=cam.projmatrix.numpy(), campos=cam.campos.numpy(), tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
              kernel_size=cam.kernel_size, subpixel_offset=None, min_depth=cam.min_depth, max_depth=cam.max_depth,
              sh_degree=sc.sh_degree)
    st = o.state()
    print("oracle forward %.1f s, R = %d" % (time.time() - t0, st["R"]))
    m2, co = np.ascontiguousarray(st["means2D"]), np.ascontiguousarray(st["conic_opacity"])
    pl, rg = np.ascontiguousarray(st["point_list"]), np.ascontiguousarray(st["ranges"])
    print("%-10s %-6s | %10s %12s %8s | %12s %8s %12s | %10s" % ("block", "batch", "cand/warp", "blend-exec", "lanes", "cursor-blend", "lanes", "cursor-steps", "R_eff"))
    for (bw, bh) in ((8, 4), (4, 8), (16, 2), (8, 8), (4, 4)):
        for batch in (256, 128):
            out = np.zeros(9, np.int64)
            t0 = time.time()
            lib.simulate(cam.W, cam.H, bw, bh, batch, 1, m2.ctypes.data, co.ctypes.data, pl.ctypes.data, rg.ctypes.data, out.ctypes.data)
            lanes = bw * bh
            print("%2dx%-7d %-6d | %10.2fM %11.2fM %7.1f%% | %11.2fM %7.1f%% %11.2fM | %9.2fM   (%.0f s; blends %.1fM, listed %.2fM)"
                  % (bw, bh, batch, out[0] / 1e6, out[1] / 1e6, 100.0 * out[2] / max(1, out[1] * lanes), out[3] / 1e6,
                     100.0 * out[4] / max(1, out[3] * lanes), out[5] / 1e6, out[7] / 1e6, time.time() - t0, out[6] / 1e6, out[8] / 1e6))
    print()
    print("lane-private cursors, faithful rounds (walk steps = per round the longest walk of any lane):")
    print("%-10s %-6s | %12s %8s | %12s %8s | %10s" % ("block", "batch", "walk-steps", "lanes", "blend-rounds", "lanes", "blends"))
    for (bw, bh) in ((8, 4), (4, 8), (4, 4)):
        for batch in (256,):
            out = np.zeros(5, np.int64)
            lib.simulate_cursor(cam.W, cam.H, bw, bh, batch, m2.ctypes.data, co.ctypes.data, pl.ctypes.data, rg.ctypes.data, out.ctypes.data)
            lanes = bw * bh
            print("%2dx%-7d %-6d | %11.2fM %7.1f%% | %11.2fM %7.1f%% | %9.1fM"
                  % (bw, bh, batch, out[0] / 1e6, 100.0 * out[3] / max(1, out[0] * lanes), out[1] / 1e6,
                     100.0 * out[2] / max(1, out[1] * lanes), out[4] / 1e6))
    print()
    print("broadcast test + per-lane FIFO of depth Q in front of the blend body (8x4 blocks, batches of 256):")
    for Q in (1, 2, 4, 8, 16, 64):
        out = np.zeros(3, np.int64)
        lib.simulate_queue(cam.W, cam.H, 8, 4, 256, Q, m2.ctypes.data, co.ctypes.data, pl.ctypes.data, rg.ctypes.data, out.ctypes.data)
        print("Q = %-3d blend rounds %6.2fM, lanes %5.1f%%, blends %.1fM" % (Q, out[0] / 1e6, 100.0 * out[1] / max(1, out[0] * 32), out[2] / 1e6))


if __name__ == "__main__":
    main()
