"""Build libmerlin_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python ppo-2dgrid_b200/build.py [--force]

The .so lands in ppo-2dgrid_b200/lib/ (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OUT = os.path.join(HERE, "lib", "libmerlin_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-extended-lambda",
    "-shared", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "-I" + INCLUDE, "-I" + CSRC,
    "--threads", "0",   # the translation units compile side by side (the step kernels are split over three of them)
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(OUT):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps)


def build_library(force=False, verbose=False):
    if not force and not stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + sources()
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
