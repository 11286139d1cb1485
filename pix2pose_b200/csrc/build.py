"""Builds libpix2pose_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch dependency)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libpix2pose_b200.so")
SOURCES = ["engine.cu", "pnp_ransac.cu", "pipeline.cu", "depth.cu", "capi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "-Xcompiler", "-Wall", "-cudart", "static"]
# The EPnP/RANSAC kernels restate OpenCV's double-precision arithmetic operation by operation (csrc/epnp_core.cuh):
# no mul+add contraction there, explicit fma() only where the original has one.
EXTRA_FLAGS = {"pnp_ransac.cu": ["-fmad=false"]}


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    srcs = [os.path.join(HERE, s) for s in SOURCES]
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(PKG), "include", "pix2pose_b200.h"))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(deps):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    for s in srcs:
        o = os.path.join(HERE, os.path.basename(s)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + EXTRA_FLAGS.get(os.path.basename(s), []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        subprocess.check_call(cmd)
        objs.append(o)
    cmd = [nvcc, "-shared", "-cudart", "static", "-o", LIB] + objs
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
