"""Builds libmgb200.so (the C-ABI library of include/mgb200.h) in-tree with nvcc for sm_100a.

    python -m multigrid_dolfinx_b200.build          # or: from multigrid_dolfinx_b200.build import build

The .so is git-ignored but travels to the GPU box with the gpurun snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmgb200.so")
SOURCES = ["mgb_engine.cu", "mgb_setup.cpp"]
HEADERS = ["mgb_kernels.cuh", "mgb_internal.h", "mgb_types.cuh", "mgb_dist.cuh", "mgb_synth.cuh", "mgb_code.cuh",
           os.path.join("..", "..", "include", "mgb200.h")]


def nvcc_path():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(os.path.join(CSRC, f)) <= t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path(), "-O3", "-std=c++17", "--split-compile", "0", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC,-fopenmp,-O3", "-shared", "-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
