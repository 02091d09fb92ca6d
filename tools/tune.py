#!/usr/bin/env python
"""Tuning aid: build variants of the CUDA library with extra -D flags (here, on the CPU box) and time them on a GPU box.

    python tools/tune.py build  name1:DEF_A=1,DEF_B=2  name2:DEF_C=3 ...      # -> ex4dgs_b200/libex4dgs_raster_<name>.so
    python tools/tune.py run [--steps 100] name1 name2 ...                     # on the GPU box; 'base' = the product library

`run` executes `bench.py` (value leg + stage timers only) once per library and prints one table row per variant.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    mode, args = sys.argv[1], sys.argv[2:]
    if mode == "build":
        from ex4dgs_b200 import build as b
        for spec in args:
            name, _, defs = spec.partition(":")
            print(b.build_variant(name, [d for d in defs.split(",") if d]))
        return
    steps = 100
    if args and args[0] == "--steps":
        steps, args = int(args[1]), args[2:]
    rows = []
    for name in args:
        env = dict(os.environ)
        if name != "base":
            env["EX4DGS_LIB"] = os.path.join(ROOT, "ex4dgs_b200", "libex4dgs_raster_%s.so" % name)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(steps), "--no-cpu-baseline",
                            "--no-train-iter", "--no-clocks", "--value-only"], env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            st = d.get("stage_ms") or {}
            rows.append((name, d["value"], d["ms_per_step"], st))
            print("%-14s %7.1f f/s  %.3f ms | %s" % (name, d["value"], d["ms_per_step"],
                                                    "  ".join("%s %.3f" % (k[:14], v) for k, v in st.items())), flush=True)
        except Exception as e:     # noqa: BLE001
            print(name, "FAILED", e, r.stderr[-2000:], flush=True)


if __name__ == "__main__":
    main()
