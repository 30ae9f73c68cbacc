"""Development aid: run (a selection of) the GPU parity tests against libkdnb_simt.so — the kernel sources executed on
the CPU, one fiber per CUDA thread (see cuda_runtime.h in this directory) — e.g.

    python tests/devtools/simt/run.py -k "build_padded or walk_ring" -x -q

Everything after the script name goes to pytest.  Sizes above ~20k particles take minutes.  This never touches the
product: it re-points the ctypes loader of THIS process at the emulated library."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

os.environ["KDNB_NO_GRAPH"] = "1"   # no CUDA graphs in the shim: plain launches
import build as simt_build  # noqa: E402

lib = simt_build.build()
from multilanguagekdtree_b200 import _lib  # noqa: E402

_lib.LIB_PATH = lib
import pytest  # noqa: E402

sys.exit(pytest.main(["-m", "gpu", os.path.join(ROOT, "tests", "test_gpu_parity.py"), *sys.argv[1:]]))
