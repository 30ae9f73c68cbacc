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
             "or test_kick_drift_bit_exact or test_simple_sim_trajectory and 1000-100 or test_quickstat_small_test_kat_gpu "
             "or test_degenerate_geometries_tree_and_walk and grid and 7 or test_one_context_walks_different_particle_counts")


@pytest.mark.timeout(900)
def test_kernel_sources_under_the_simt_model_match_the_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "devtools", "simt", "run.py"), "-q", "-x", "-k", SELECTION,
                        "-p", "no:cacheprovider"], capture_output=True, text=True, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail


@pytest.mark.timeout(900)
@pytest.mark.parametrize("ipc", ["1", "0"])
def test_two_ranks_under_the_simt_model_are_bit_identical_to_one(ipc):
    """World size 2 of the multi-GPU path on the CPU: the ranks are threads of one process, every rank's kernels run
    under the SIMT model, NCCL is tests/devtools/simt/fake_nccl.cpp and cudaIpc handles are plain pointers
    (tests/devtools/simt/multirank.py).  ipc=1: accelerations exchanged by peer stores from inside the walk kernel +
    flag handshake + wait kernel; ipc=0: peer mapping "unavailable", the library falls back on ncclAllGather.  Either way
    replicated build + sharded walk + exchange + kick/drift, and the sharded host-buffer calls, must give every rank
    the bits of a single-rank run."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "devtools", "simt", "multirank.py"), "2", "3000", "3"],
                       capture_output=True, text=True, cwd=ROOT, env={**os.environ, "SIMT_IPC": ipc})
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and "multirank ok: world=2" in r.stdout, tail
    assert ("peer stores" in r.stdout) == (ipc == "1"), tail


@pytest.mark.timeout(600)
def test_ranks_with_empty_shards_under_the_simt_model():
    """Fewer particles than 64 x (world - 1): some ranks have nothing to walk.  In peer mode they must still raise their
    flag (one idle CTA) or every rank waits for it until the wait kernel's timeout, every step — found by this model."""
    for world, n in (("2", "33"), ("4", "100")):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "devtools", "simt", "multirank.py"), world, n, "2"],
                           capture_output=True, text=True, cwd=ROOT, timeout=300)
        assert r.returncode == 0 and f"multirank ok: world={world}" in r.stdout, (r.stdout + r.stderr)[-2000:]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", ["2", "4"])
def test_sharded_tree_build_under_the_simt_model(world):
    """The sharded build of the multi-GPU path (build.cu: every rank builds one subtree below the top log2(world) levels,
    pushes its node records and tree order to the peers, gathers the foreign particle copies itself), forced at a small
    size: every rank's tree — node records and tree order —, accelerations and trajectory must be bit-identical to a
    single-rank run (tests/devtools/simt/multirank.py compares all of them)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "devtools", "simt", "multirank.py"), world, "9000", "2"],
                       capture_output=True, text=True, cwd=ROOT, env={**os.environ, "KDNB_SHARD_BUILD": "1", "KDNB_SORT_SPLIT": "0"})
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and f"multirank ok: world={world}" in r.stdout, tail
    # the sharded path really ran: it adds four launches per step (push, wait, foreign gather, top levels' sums)
    import re
    launches = float(re.search(r"([0-9.]+) kernel launches per step", r.stdout).group(1))
    r0 = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "devtools", "simt", "multirank.py"), world, "9000", "2"],
                        capture_output=True, text=True, cwd=ROOT, env={**os.environ, "KDNB_SHARD_BUILD": "0", "KDNB_SORT_SPLIT": "0"})
    base = float(re.search(r"([0-9.]+) kernel launches per step", r0.stdout).group(1))
    assert launches > base + 3.5, (launches, base)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", ["2", "3"])
def test_split_sort_under_the_simt_model(world):
    """The split sort of the multi-GPU path (sort.cu: a rank sorts only its share of the non-flat dimensions, exports the
    lists and fetches the others from a peer), forced at a small size together with the sharded build: trees,
    accelerations and trajectories bit-identical to a single-rank run on every rank."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "devtools", "simt", "multirank.py"), world, "9000", "2"],
                       capture_output=True, text=True, cwd=ROOT, env={**os.environ, "KDNB_SHARD_BUILD": "1", "KDNB_SORT_SPLIT": "1"})
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and f"multirank ok: world={world}" in r.stdout, tail
