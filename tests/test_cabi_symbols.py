"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol include/kdnb.h declares,
its pure functions agree with the oracle, and it refuses to run (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import multilanguagekdtree_b200 as kd
from multilanguagekdtree_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "kdnb.h")).read()
    declared = set(re.findall(r"\b(kdnb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.load()
    for s in declared:
        assert hasattr(L, s), s
    assert L.kdnb_version() == 100


def test_struct_layouts_match_header():
    assert _lib.PARTICLE.itemsize == 64 and _lib.NODE.itemsize == 88
    assert C.sizeof(_lib.Config) == 32
    assert _lib.NODE.fields["cm"][1] == 40 and _lib.NODE.fields["right"][1] == 80


def test_nodes_needed_matches_oracle(orc):
    for mp in (4, 7, 8, 16, 32):
        for n in list(range(1, 400)) + [1001, 5001, 100001, 1000001, 10000001, 100000001]:
            assert kd.nodes_needed_for_particles(n, mp) == orc.nodes_needed_for_particles(n, mp)
    assert len(kd.allocate_node_vec(1000001)) == 524287


def test_ic_generators_match_oracle(orc):
    a, b = kd.circular_orbits(5000, seed=99), orc.circular_orbits(5000, seed=99)
    # same seeded angle stream; numpy and glibc cos/sin may differ in the last ulp
    assert np.allclose(a["p"], b["p"], rtol=0, atol=1e-15) and np.allclose(a["v"], b["v"], rtol=0, atol=1e-14)
    assert np.array_equal(a["m"], b["m"]) and np.array_equal(a["r"], b["r"])
    assert kd.two_bodies().tobytes() == orc.two_bodies().tobytes()


def test_no_cpu_fallback_without_device():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = os.path.exists("/dev/nvidia0")
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(kd.KdnbError, match="no CUDA device"):
        kd.KDTreeSim()
    with pytest.raises(kd.KdnbError):
        kd.simple_sim(kd.circular_orbits(10), 1e-3, 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multilanguagekdtree_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "liboracle", "okd_", "kdtree_oracle"):
                    assert needle not in src, (f, needle)
