"""In-tree build of the CUDA library (sm_100a only) with explicit nvcc commands.

    python -m ex4dgs_b200.build            # build if stale
    python -m ex4dgs_b200.build --force

Produces ex4dgs_b200/libex4dgs_raster.so (git-ignored, shipped to the GPU box by gpurun).
No torch dependency: the library is plain CUDA runtime + CUB behind a C ABI.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libex4dgs_raster.so")
SOURCES = ["api.cu", "preprocess.cu", "binning.cu", "render_fwd.cu", "render_bwd.cu", "frontend.cu", "loss.cu", "optim.cu", "stats.cu", "compact.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "ex4dgs_raster.h")]
# No -use_fast_math: IEEE sqrt/div and libdevice expf are part of the parity contract.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "--compiler-options", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
              "-diag-suppress", "177"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(src: str, force: bool, verbose: bool) -> str:
    s = os.path.join(CSRC, src)
    o = os.path.join(OBJ, src + ".o")
    if force or _stale(o, [s] + HEADERS):
        cmd = ["nvcc", "-c", s, "-o", o] + NVCC_FLAGS
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(o + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, log))
        if verbose:
            print(log)
    return o


def build_variant(name: str, defines) -> str:
    """Experiment helper: build libex4dgs_raster_<name>.so with extra -D flags (selected at run time
    with EX4DGS_LIB=<path>); the product always loads libex4dgs_raster.so."""
    objdir = os.path.join(HERE, "_obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src in SOURCES:
        o = os.path.join(objdir, src + ".o")
        subprocess.check_call(["nvcc", "-c", os.path.join(CSRC, src), "-o", o] + NVCC_FLAGS + ["-D" + d for d in defines],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        objs.append(o)
    lib = os.path.join(HERE, "libex4dgs_raster_%s.so" % name)
    subprocess.check_call(["nvcc", "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    if not os.path.isdir(CSRC):
        raise RuntimeError("csrc missing")
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = [os.path.join(CSRC, s) for s in srcs] + HEADERS
    if not force and not _stale(LIB, deps):
        return LIB
    import shutil
    if shutil.which("nvcc") is None:
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("nvcc not found and %s is missing" % LIB)
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    cmd = ["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
