#!/usr/bin/env python
"""Profiling aid: a few frames (fwd+bwd at the GaussianRasterizer boundary, workload C3 by default) with only the LAST
ones inside cudaProfilerStart/Stop, for

    ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/launches.csv python tools/one_frame.py
    ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:render_ -o gpurun_out/prof python tools/one_frame.py
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ex4dgs_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--impl", default="ours")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    if args.impl == "ours":
        import ex4dgs_b200 as mod
    else:
        mod = bench.load_reference()
    fr = bench.Frame(mod, synth.make_config(args.workload), dev, 0, impl=args.impl)
    for _ in range(args.warmup):
        fr.step_device()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(args.frames):
        fr.step_device()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
