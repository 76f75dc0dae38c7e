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
SOURCES = ["mgb_engine.cu", "mgb_devsetup.cu", "mgb_setup.cpp"]
HEADERS = ["mgb_kernels.cuh", "mgb_internal.h", "mgb_types.cuh", "mgb_dist.cuh", "mgb_synth.cuh", "mgb_code.cuh", "mgb_devsetup.h",
           os.path.join("..", "..", "include", "mgb200.h")]
# what each translation unit includes (a header that changes only recompiles the units that see it)
DEPENDS = {"mgb_engine.cu": HEADERS,
           "mgb_devsetup.cu": ["mgb_devsetup.h"],
           "mgb_setup.cpp": ["mgb_internal.h", os.path.join("..", "..", "include", "mgb200.h")]}
OBJDIR = os.path.join(HERE, "_build")
SPLIT_COMPILE = 4
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC,-fopenmp,-O3"]


def nvcc_path():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _newer(path, t):
    return os.path.getmtime(path) > t


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return not any(_newer(os.path.join(CSRC, f), t) for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """One object per source under _build/ (git-ignored), compiled in parallel and only when the source or a header it
    includes is newer; then one link.  Every object is sm_100a SASS + compute_100a PTX with -lineinfo."""
    if not force and up_to_date():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = nvcc_path()
    jobs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or any(
            _newer(os.path.join(CSRC, f), os.path.getmtime(obj)) for f in [src] + DEPENDS[src])
        if stale:
            # --split-compile: the optimiser works on the translation unit in pieces, and the PTX (hence ptxas's register allocation
            # and instruction counts, +-3 %) depends on how many.  With "0 = one piece per core" the library followed the build box's
            # core count, so the count is fixed at 4, the value the shipped, GPU-tested library of round 2 was built with.  Even so,
            # repeated builds of mgb_engine.cu on one 8-core box gave EITHER the 4-piece PTX (33.5 MB) or exactly the PTX that 8 pieces
            # give (41.6 MB), not tied to anything in the launching environment; ptxas is deterministic for a given PTX.  The
            # library shipped at the end of round 2 (final GPU test run, smoke) is the 4-piece variant.  What parity depends on -- no fused multiply-add in any row-sum kernel -- holds
            # in both and is checked on whatever library is present by tests/test_sass.py.
            cmd = [nvcc] + FLAGS + (["--split-compile", str(SPLIT_COMPILE)] if src.endswith(".cu") else []) + \
                  (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
            jobs.append((src, subprocess.Popen(cmd)))
    failed = [src for src, p in jobs if p.wait() != 0]
    if failed:
        raise RuntimeError(f"nvcc failed on {failed}")
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC,-fopenmp",
                           "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
