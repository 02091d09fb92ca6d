"""Diagnostic (variant library built with -DEX_BWD_HIST=1): distribution of the number of lanes / pixels that contribute
per processed (8x8 block, splat) entry of the backward compositing kernel at C3."""
import ctypes as C
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ex4dgs_b200 import synth, _lib  # noqa: E402
import ex4dgs_b200 as mod  # noqa: E402
dev = torch.device("cuda", 0)
fr = bench.Frame(mod, synth.make_config("C3"), dev, 0)
lib = _lib.load()
out = (C.c_ulonglong * 66)()
fr.step_device(); torch.cuda.synchronize()
lib.ex4dgs_debug_bwd_hist(out)
fr.step_device(); torch.cuda.synchronize()
lib.ex4dgs_debug_bwd_hist(out)
h = list(out)
tot = sum(h[:33])
print("entries", tot)
cum = 0
for i in range(33):
    cum += h[i]
    print("lanes %2d: %6.2f %%  cum %6.2f %%   | pixels %2d-%2d: %6.2f %%" % (i, 100.0 * h[i] / tot, 100.0 * cum / tot, 2 * i, 2 * i + 1, 100.0 * h[33 + i] / tot))
