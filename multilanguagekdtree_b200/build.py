"""Build libkdnb.so (hand-written sm_100a CUDA behind the C ABI of include/kdnb.h) and the kdtree-sim CLI, in-tree.

    python -m multilanguagekdtree_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  -fmad=false: the reference (rustc) never contracts a*b+c, and the
acceptance test / kick-drift / node sums must round exactly like it; FMAs are written explicitly where allowed.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkdnb.so")
CLI = os.path.join(HERE, "kdtree-sim")
CU = ["kdnb_api.cu", "sort.cu", "build.cu", "walk.cu", "kick.cu", "select.cu", "peak.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.exists(s) and os.path.getmtime(s) <= t for s in sources)  # (a listed file that is gone: rebuild)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, f) for f in CU]
    deps = srcs + [os.path.join(CSRC, f) for f in ("common.cuh", "ctx.cuh", "walk2.cuh")] + [os.path.join(HERE, "..", "include", "kdnb.h")]
    extra = os.environ.get("KDNB_NVCC_EXTRA", "").split()   # compile-time experiment knobs, e.g. -DKDNB_BOT_CAP=1024
    if force or extra or not _newer(LIB, deps):
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-shared", "-o", LIB, *srcs, "-ldl"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
    cli_src = os.path.join(CSRC, "kdtree_sim.cpp")
    if os.path.exists(cli_src) and (force or not _newer(CLI, [cli_src, LIB, os.path.join(CSRC, "kdnb.hpp")])):
        cmd = ["g++", "-O2", "-std=c++17", "-o", CLI, cli_src, "-L" + HERE, "-lkdnb", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
