"""Development aid: compile the kernel sources of multilanguagekdtree_b200/csrc with g++ against the SIMT shim in this
directory (cuda_runtime.h, simt_runtime.cpp) into _build/libkdnb_simt.so — the same C ABI, every CUDA thread a fiber on
the CPU.  Not product, not a fallback: nothing under multilanguagekdtree_b200/ knows about it (see run.py)."""
import fcntl
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", "..", ".."))
CSRC = os.path.join(ROOT, "multilanguagekdtree_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libkdnb_simt.so")
CU = ["kdnb_api.cu", "sort.cu", "build.cu", "walk.cu", "kick.cu", "select.cu", "peak.cu"]


def build(verbose: bool = False) -> str:
    """One builder at a time (concurrent test processes — pytest -n — serialise on a lock file; the losers find everything
    up to date), objects and library written under temporary names and renamed into place."""
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(os.path.join(os.path.dirname(OUT), ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        return _build_locked(verbose)


def _build_locked(verbose: bool) -> str:
    objs, rebuilt = [], False
    flags = ["-std=c++17", "-O1", "-g", "-fPIC", "-fopenmp", "-ffp-contract=off", "-mfma", "-Wno-unknown-pragmas",
             "-Wno-attributes", "-I", HERE, "-I", CSRC, *os.environ.get("KDNB_NVCC_EXTRA", "").split()]
    force = bool(os.environ.get("KDNB_NVCC_EXTRA"))
    for f in CU + ["simt_runtime.cpp"]:
        src = os.path.join(CSRC, f) if f.endswith(".cu") else os.path.join(HERE, f)
        obj = os.path.join(HERE, "_build", f.replace(".cu", ".o").replace(".cpp", ".o"))
        deps = [src, os.path.join(HERE, "cuda_runtime.h")] + [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".hpp"))]
        if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
            objs.append(obj)
            continue
        cmd = ["g++", *flags, "-x", "c++", "-c", src, "-o", obj + ".tmp"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed on {f}:\n" + r.stderr[-6000:])
        os.replace(obj + ".tmp", obj)
        rebuilt = True
        if verbose and r.stderr:
            print(r.stderr[-2000:])
        objs.append(obj)
    if not rebuilt and os.path.exists(OUT) and all(os.path.getmtime(o) <= os.path.getmtime(OUT) for o in objs):
        return OUT
    r = subprocess.run(["g++", "-shared", "-fopenmp", "-o", OUT + ".tmp", *objs, "-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
