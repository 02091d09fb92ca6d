#!/usr/bin/env python
"""Turn one `ncu --set full` capture of the two compositing kernels (taken with tools/one_frame.py on THIS tree) into
the tracked artefacts under profiles/:

    python tools/make_profile_artifacts.py gpurun_out/r2X_full.ncu-rep r2X

  profiles/<tag>_ncu_full_render_fwd_bwd.csv   raw page of the capture (every metric of both kernels)
  profiles/traffic.json                        DRAM bytes and warp instructions per launch + the hash of the kernel
                                               sources they belong to (bench.py only uses them while the hash matches)
  profiles/<tag>_sass_render_fwd.txt, ..._bwd.txt   opcode histograms of the built library's compositing kernels
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    import bench
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    out_csv = os.path.join(ROOT, "profiles", "%s_ncu_full_render_fwd_bwd.csv" % tag)
    open(out_csv, "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    tj = {"kernel_source_sha16": bench.kernel_source_hash(), "capture": os.path.basename(out_csv)}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        key = "render_fwd_kernel" if "render_fwd" in name else "render_bwd_kernel" if "render_bwd" in name else None
        if key is None or key in tj:
            continue
        g = lambda m: float(r[hdr.index(m)].replace(",", ""))     # noqa: E731
        unit = rows[1][hdr.index("dram__bytes_read.sum")]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        tj[key] = {"dram_bytes_per_launch": int((g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) * scale),
                   "inst_executed_per_launch": int(g("smsp__inst_executed.sum")),
                   "gpu_time_us": g("gpu__time_duration.sum"),
                   "kernel": name,
                   "source": "ncu --set full, %s (dram__bytes_read.sum + dram__bytes_write.sum of one launch at C3, tile-cull on)" % os.path.basename(out_csv)}
    json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(tj, indent=1))
    # SASS opcode histograms
    lib = os.path.join(ROOT, "ex4dgs_b200", "libex4dgs_raster.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, per = None, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            per.setdefault(cur, {}).setdefault(m.group(2), 0)
            per[cur][m.group(2)] += 1
    for which in ("render_fwd", "render_bwd"):
        with open(os.path.join(ROOT, "profiles", "%s_sass_%s.txt" % (tag, which)), "w") as f:
            f.write("# cuobjdump -sass ex4dgs_b200/libex4dgs_raster.so : static opcode counts per instantiation of %s_kernel\n" % which)
            f.write("# packed FP32 = FFMA2 / FMUL2 / FADD2; TMA bulk copies = UBLKCP; cp.async = LDGSTS; warp reductions = REDG / SHFL.BFLY\n")
            for fn, ops in per.items():
                if which not in fn:
                    continue
                tot = sum(ops.values())
                f.write("\n%s   (%d instructions)\n" % (fn, tot))
                grp = {}
                for op, n in ops.items():
                    grp[op.split(".")[0]] = grp.get(op.split(".")[0], 0) + n
                for op, n in sorted(grp.items(), key=lambda kv: -kv[1]):
                    f.write("  %-10s %5d\n" % (op, n))
                for op in ("UBLKCP.S.G", "LDGSTS.E.BYPASS.128", "REDG.E.ADD.F32.FTZ.RN.STRONG.GPU", "SHFL.BFLY", "MUFU.EX2", "MUFU.RCP"):
                    if op in ops:
                        f.write("  [%s x %d]\n" % (op, ops[op]))


if __name__ == "__main__":
    main()
