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


# ---- the Rust binding (source only: no rustc in this image) is at least kept in step with the header -----------

def _c_prototypes():
    header = open(os.path.join(ROOT, "include", "kdnb.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = {}
    for name, args in re.findall(r"\b(kdnb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", header):
        args = args.strip()
        protos[name] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def _rust_sources(crate):
    out = {}
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rust", crate)):
        for f in files:
            if f.endswith(".rs"):
                out[os.path.relpath(os.path.join(dirpath, f), ROOT)] = open(os.path.join(dirpath, f)).read()
    return out


def test_rust_sys_crate_declares_every_entry_point_with_the_same_arity():
    protos = _c_prototypes()
    src = _rust_sources("kdnb-sys")["rust/kdnb-sys/src/lib.rs"]
    rust = {name: (0 if not args.strip() else args.count(":"))
            for name, args in re.findall(r"pub fn (kdnb_[a-z0-9_]+)\(([^)]*)\)", src)}
    assert set(rust) == set(protos), set(rust) ^ set(protos)
    for name, n in protos.items():
        assert rust[name] == n, (name, rust[name], n)
    # records: same field order as the C structs
    for struct, fields in (("kdnb_particle", ["p", "v", "r", "m"]),
                           ("kdnb_node", ["kind", "split_dim", "num_parts", "leaf_first", "split_val", "m", "cm", "size", "left", "right"]),
                           ("kdnb_config", ["struct_size", "device", "max_parts", "layout", "theta", "flags", "reserved"])):
        body = re.search(r"pub struct %s \{(.*?)\}" % struct, src, flags=re.S).group(1)
        assert re.findall(r"pub (\w+):", body) == fields, struct


def test_rust_host_crate_mirrors_the_reference_api_and_only_calls_declared_symbols():
    protos = _c_prototypes()
    files = _rust_sources("rust_kdtree_nbody")
    assert {"rust/rust_kdtree_nbody/src/lib.rs", "rust/rust_kdtree_nbody/src/main.rs", "rust/rust_kdtree_nbody/src/gpu.rs",
            "rust/rust_kdtree_nbody/src/array_kd_tree.rs", "rust/rust_kdtree_nbody/src/array_particle.rs",
            "rust/rust_kdtree_nbody/src/quickstat.rs"} <= set(files)
    for path, src in files.items():
        for sym in re.findall(r"sys::(kdnb_[a-z0-9_]+)\s*\(", src):
            assert sym in protos, (path, sym)
        code = re.sub(r"//.*", "", src)
        code = re.sub(r'"(?:[^"\\]|\\.)*"', '""', code)
        code = re.sub(r"'(?:[^'\\]|\\.)'", "' '", code)
        for a, b in ("{}", "()", "[]"):
            assert code.count(a) == code.count(b), (path, a, code.count(a), code.count(b))
    # the reference crate's modules and the `pub` items of the hot path, same names
    lib = files["rust/rust_kdtree_nbody/src/lib.rs"]
    for mod in ("array_kd_tree", "array_particle", "quickstat"):
        assert f"pub mod {mod};" in lib
    tree = files["rust/rust_kdtree_nbody/src/array_kd_tree.rs"]
    for item in ("pub const MAX_PARTS: usize = 8", "pub const THETA: f64 = 0.3", "pub enum KDTree", "pub fn leaf(",
                 "pub fn nodes_needed_for_particles(", "pub fn allocate_node_vec(", "pub fn build_tree(",
                 "pub fn build_tree_par4(", "pub fn simple_sim(bodies: &mut Vec<Particle>, dt: f64, steps: i64)",
                 "pub fn print_tree("):
        assert item in tree, item
    part = files["rust/rust_kdtree_nbody/src/array_particle.rs"]
    for item in ("#[repr(C)]", "pub struct Particle", "pub fn two_bodies() -> Vec<Particle>",
                 "pub fn circular_orbits(n: usize) -> Vec<Particle>", "pub fn calc_pp_accel("):
        assert item in part, item
    main = files["rust/rust_kdtree_nbody/src/main.rs"]
    assert '"--number"' in main and '"--steps"' in main and "1e-3" in main


def test_incremental_build_entry_point_with_the_library_already_built():
    """__graft_entry__.build() on a tree whose libkdnb.so exists takes the up-to-date check: every dependency it lists
    must exist (a stale name made the check raise instead of answering), and the call must be a no-op that succeeds."""
    from multilanguagekdtree_b200 import build as kbuild
    assert os.path.exists(kbuild.LIB)
    src = open(kbuild.__file__).read()
    for name in re.findall(r'"([A-Za-z0-9_]+\.(?:cu|cuh|hpp|cpp|h))"', src):
        cands = [os.path.join(kbuild.CSRC, name), os.path.join(kbuild.HERE, "..", "include", name)]
        assert any(os.path.exists(c) for c in cands), f"build.py names {name}, which does not exist"
    before = os.path.getmtime(kbuild.LIB)
    assert kbuild.build(force=False) == kbuild.LIB
    assert os.path.getmtime(kbuild.LIB) == before
