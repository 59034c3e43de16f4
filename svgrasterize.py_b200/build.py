"""Builds libsvgr_b200.so (the C-ABI library with the sm_100a kernels) in-tree.

    python -m svgrasterize_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  -fmad=false: the flatten and
stroke kernels reproduce the reference's float64 rounding recipes and spell
every fused multiply-add explicitly.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsvgr_b200.so")
SOURCES = ["engine.cu", "k_flatten.cu", "k_stroke.cu", "k_coverage.cu", "k_compose.cu", "k_filters.cu",
           "k_stencil_tma.cu", "k_png.cu", "encode_flat.cpp", "pathdata.cpp"]
HEADERS = ["svgr_types.h", "svgr_kernels.h", "svgr_device.cuh", os.path.join("..", "..", "include", "svgr_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall", "--cudart", "static",
]
# Geometry (flatten, stroke, coverage edge math, the planner) reproduces the reference's float64 roundings and
# must not have multiplies and adds contracted behind its back; the pixel kernels are float32 work within a
# 1e-5 tolerance and keep the default contraction (FFMA).
FMAD = {"k_compose.cu": "true", "k_filters.cu": "true", "k_stencil_tma.cu": "true", "k_png.cu": "true"}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


EXTRA = os.environ.get("SVGR_NVCC_EXTRA", "").split()


FLATTEN_SRC = os.path.join(CSRC, "_flatten.c")


def flatten_path() -> str:
    import sysconfig

    return os.path.join(HERE, "_svgr_flatten" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_flatten(force: bool = False) -> str:
    """The CPython extension that reads Scene objects into flat arrays (csrc/_flatten.c), built in-tree with gcc."""
    import sysconfig

    out = flatten_path()
    deps = [FLATTEN_SRC, os.path.join(HERE, "..", "include", "svgr_b200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cc = os.environ.get("CC") or shutil.which("gcc") or "cc"
    import numpy

    cmd = [cc, "-O2", "-shared", "-fPIC", "-Wall", "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
           FLATTEN_SRC, "-o", out]
    subprocess.run(cmd, check=True)
    return out


def build(force: bool = False, verbose: bool = False, lib: str = LIB) -> str:
    if lib == LIB:
        build_flatten(force)
    if not force and lib == LIB and not stale():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(CSRC, "build", os.path.splitext(src)[0] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, f"-fmad={FMAD.get(src, 'false')}", *EXTRA, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    cmd = [nvcc, "-shared", "--cudart", "static", "-Wno-deprecated-gpu-targets", "-o", lib, *objs]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    out = [a for a in sys.argv[1:] if a.endswith(".so")]
    print(build(force="--force" in sys.argv or bool(out), verbose="-v" in sys.argv, lib=os.path.abspath(out[0]) if out else LIB))
