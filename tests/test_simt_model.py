"""CPU tier: the kernel SOURCES (csrc/*.cu) executed under the SIMT model of tests/devtools/simt — every CUDA thread a
fiber on the CPU, same C ABI — must reproduce the oracle on small inputs: tree topology and leaf membership bit for bit
(padded and dense layouts, planar and 3-D), every accept/open decision of the walk, accelerations and a short
trajectory within 1e-12.  This checks kernel LOGIC (barrier placement, look-back protocol, list hand-offs) on machines
without a GPU; it is a development aid run in its own process, not a code path of the product (the product loads
libkdnb.so only and fails without a CUDA device: tests/test_cabi_symbols.py::test_no_cpu_fallback_without_device).
The `-m gpu` tests remain the parity tests proper."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELECTION = ("test_build_padded_ring_bit_exact and (2049 or 12293) or test_build_dense_bit_exact and 5001 "
             "or test_build_3d_and_ties_bit_exact and 3000 or test_walk_ring_acc_and_decisions and 5000 "
             "or test_walk_production_kernel_equals_counted_kernel and cube_unequal or test_walk_equal_mass_cube and 0.5 "
             "or test_kick_drift_bit_exact or test_simple_sim_trajectory and 1000-100 or test_quickstat_small_test_kat_gpu")


@pytest.mark.timeout(900)
def test_kernel_sources_under_the_simt_model_match_the_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "devtools", "simt", "run.py"), "-q", "-x", "-k", SELECTION,
                        "-p", "no:cacheprovider"], capture_output=True, text=True, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
