"""Development aid: run (a selection of) the GPU parity tests against libkdnb_simt.so — the kernel sources executed on
the CPU, one fiber per CUDA thread (see cuda_runtime.h in this directory) — e.g.

    python tests/devtools/simt/run.py -k "build_padded or walk_ring" -x -q

Everything after the script name goes to pytest.  Sizes above ~20k particles take minutes (deselect
test_walk_production_kernel_on_a_grid_of_several_waves: 300k particles).  This never touches the
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

# Tests that start a fresh process (a python child importing the package, the kdtree-sim binary) are re-pointed too, from
# the TEST side only: a sitecustomize on the children's PYTHONPATH sets the loader path, and a directory holding a
# `libkdnb.so` link to the emulated library goes first on LD_LIBRARY_PATH (kdtree-sim finds its library by RUNPATH,
# which is searched after it).  Nothing under multilanguagekdtree_b200/ knows about either.
_site = os.path.join(HERE, "_build", "site")
_libdir = os.path.join(HERE, "_build", "libdir")
os.makedirs(_site, exist_ok=True)
os.makedirs(_libdir, exist_ok=True)
# (written under process-private names and renamed into place: several of these runners may start at once, pytest -n)
_tmp = os.path.join(_site, "sitecustomize.py.%d" % os.getpid())
with open(_tmp, "w") as f:
    f.write("import sys\nsys.path.insert(0, %r)\nfrom multilanguagekdtree_b200 import _lib\n_lib.LIB_PATH = %r\n" % (ROOT, lib))
os.replace(_tmp, os.path.join(_site, "sitecustomize.py"))
_link = os.path.join(_libdir, "libkdnb.so")
_tmp = _link + ".%d" % os.getpid()
os.symlink(lib, _tmp)
os.replace(_tmp, _link)
os.environ["PYTHONPATH"] = _site + os.pathsep + os.environ.get("PYTHONPATH", "")
os.environ["LD_LIBRARY_PATH"] = _libdir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", "")
import pytest  # noqa: E402

sys.exit(pytest.main(["-m", "gpu", os.path.join(ROOT, "tests", "test_gpu_parity.py"), *sys.argv[1:]]))
