#!/usr/bin/env python
"""Build the UNMODIFIED reference rasterizer extension into oracle/_ref/ (test infrastructure).

This is the "real reference" leg of the oracle (see oracle/README.md): the reference CUDA
extension `diff_gaussian_rasterization_df._C` is compiled *from the sources where they lie*
under /root/reference/submodules/diff_gaussian_rasterization_df (nothing is copied into the
repository history; oracle/_ref/ is git-ignored but travels to the GPU box with gpurun).

It does NOT run the reference's own build system (setup.py / CMakeLists.txt): the five
translation units are compiled with explicit nvcc / g++ commands below.  The only deviation
from a stock build is `-include cstdint`, required because rasterizer_impl.h:24,40-61 uses
std::uintptr_t / uint32_t / uint64_t without including <cstdint> (fails on gcc 13).

The result is laid out like `pip install --target oracle/_ref` would lay it out:
    oracle/_ref/diff_gaussian_rasterization_df/__init__.py   (installed copy of the binding)
    oracle/_ref/diff_gaussian_rasterization_df/_C.so         (sm_100 cubins)
so `sys.path.insert(0, "oracle/_ref")` exposes the reference's own public API.  Only tests/,
bench.py (--impl reference) and __graft_entry__.smoke() may import it.

The reference path is CUDA-only (rasterize_points.cu:80 hard-codes torch::kCUDA), so the
built module can only be *executed* on the GPU box; here it is just compiled.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/submodules/diff_gaussian_rasterization_df"
OUT = os.path.join(HERE, "_ref", "diff_gaussian_rasterization_df")
OBJ = os.path.join(HERE, "_ref", "obj")


def run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build(force=False):
    so = os.path.join(OUT, "_C.so")
    if not os.path.isdir(REF):
        print("reference sources not present (GPU box?) - using prebuilt", so)
        return os.path.exists(so)
    srcs = [os.path.join(REF, p) for p in (
        "cuda_rasterizer/rasterizer_impl.cu", "cuda_rasterizer/forward.cu",
        "cuda_rasterizer/backward.cu", "rasterize_points.cu", "ext.cpp")]
    if os.path.exists(so) and not force:
        newest = max(os.path.getmtime(s) for s in srcs + [os.path.abspath(__file__)])
        if os.path.getmtime(so) >= newest:
            install_callers()
            return True
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    incs = []
    for p in ce.include_paths() + [sysconfig.get_paths()["include"],
                                   os.path.join(REF, "third_party/glm"), "/usr/local/cuda/include"]:
        incs += ["-I", p]
    defs = ["-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s) + ".o")
        if s.endswith(".cu"):
            run(["nvcc", "-c", s, "-o", o, "-std=c++17", "-include", "cstdint",
                 "-gencode", "arch=compute_100,code=sm_100", "--compiler-options", "-fPIC",
                 "-w", "--expt-relaxed-constexpr"] + defs + incs)
        else:
            run(["g++", "-c", s, "-o", o, "-std=c++17", "-fPIC", "-O2", "-w"] + defs + incs)
        objs.append(o)
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    run(["g++", "-shared", "-o", so] + objs + [
        "-L", libdir, "-L", "/usr/local/cuda/lib64", "-Wl,-rpath," + libdir,
        "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"])
    # "install" the binding next to the extension, as pip --target would (git-ignored dir).
    shutil.copyfile(os.path.join(REF, "diff_gaussian_rasterization_df", "__init__.py"),
                    os.path.join(OUT, "__init__.py"))
    install_callers()
    return True


def install_callers():
    """Install (git-ignored, like the extension) the reference's ONLY caller of the rasterizer,
    gaussian_renderer/__init__.py, and its one pure-PyTorch import (utils/sh_utils.py), so that the
    GPU box can run the unmodified render() against this repository's drop-in package
    (tests/test_gpu_dropin.py); plus utils/loss_utils.py (and its import utils/graphics_utils.py),
    the loss block of train.py:144-151, so that `bench.py --impl reference` can time the reference's
    own training-step loss next to the fused one (row N2); plus scene/c_gaussian_model.py and the utils modules it
    imports, so that the reference arm's training iteration runs the reference's OWN statistics methods
    (mark_prune_stats, add_densification_stats, add_l1_ssim_stats, prune_nan_points) next to the fused kernel."""
    root = "/root/reference"
    dst = os.path.join(HERE, "_ref", "callers")
    for rel in ("gaussian_renderer/__init__.py", "utils/sh_utils.py", "utils/loss_utils.py", "utils/graphics_utils.py",
                "scene/c_gaussian_model.py", "utils/general_utils.py", "utils/system_utils.py", "utils/interpolations.py"):
        os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
        shutil.copyfile(os.path.join(root, rel), os.path.join(dst, rel))
    open(os.path.join(dst, "utils", "__init__.py"), "a").close()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("reference extension:", "built" if ok else "unavailable")
