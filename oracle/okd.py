"""ctypes binding of the CPU oracle (oracle/kdtree_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under multilanguagekdtree_b200/ may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

PARTICLE = np.dtype([("p", "<f8", (3,)), ("v", "<f8", (3,)), ("r", "<f8"), ("m", "<f8")], align=True)
assert PARTICLE.itemsize == 64

LAYOUT_PADDED, LAYOUT_DENSE = 0, 1
ORDER_FAITHFUL, ORDER_CANONICAL = 0, 1
U64MAX = np.uint64(0xFFFFFFFFFFFFFFFF)


def _node_dtype(cap: int) -> np.dtype:
    # okd_node: u64 tag + union{ leaf{u64 n; u64 parts[cap]} | in{u64 sd; f64 sv, m, cm[3], size; u64 l, r} }
    union_bytes = max(8 * (1 + cap), 72)
    return np.dtype(
        {
            "names": ["is_internal", "num_parts", "leaf_parts", "split_dim", "split_val", "m", "cm", "size", "left", "right"],
            "formats": ["<u8", "<u8", ("<u8", (cap,)), "<u8", "<f8", "<f8", ("<f8", (3,)), "<f8", "<u8", "<u8"],
            "offsets": [0, 8, 16, 8, 16, 24, 32, 56, 64, 72],
            "itemsize": 8 + union_bytes,
        }
    )


WALK_COUNTS = np.dtype([("node_visits", "<u8"), ("accepts", "<u8"), ("leaf_visits", "<u8"), ("pp", "<u8")])


def build(force: bool = False) -> None:
    """Compile the oracle (and, when /root/reference is present, oracle/_ref)."""
    lib = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "kdtree_oracle.c")
    if force or not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/Parallel/CppVersion") and (
        force or not os.path.exists(os.path.join(_HERE, "_ref", "kdtree-sim-cpp"))
    ):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


class Oracle:
    def __init__(self, leaf_cap: int = 8):
        assert leaf_cap in (8, 32)
        name = "liboracle.so" if leaf_cap == 8 else "liboracle32.so"
        path = os.path.join(_HERE, "_build", name)
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        self.leaf_cap = leaf_cap
        self.NODE = _node_dtype(leaf_cap)
        L = self.lib
        u64, i64, f64, vp, i32 = C.c_uint64, C.c_int64, C.c_double, C.c_void_p, C.c_int
        L.okd_rng_next.restype = u64
        L.okd_rng_next.argtypes = [vp]
        L.okd_two_bodies.argtypes = [vp]
        L.okd_circular_orbits.argtypes = [u64, u64, vp]
        L.okd_nodes_needed_for_particles.restype = u64
        L.okd_nodes_needed_for_particles.argtypes = [u64, u64]
        L.okd_fill_default_nodes.argtypes = [vp, u64]
        L.okd_quickstat_index_f64.argtypes = [vp, u64, u64, vp, vp]
        L.okd_build_tree.restype = u64
        L.okd_build_tree.argtypes = [vp, u64, u64, vp, u64, vp, u64, u64, vp]
        L.okd_build_tree_par4.argtypes = [vp, u64, u64, vp, vp, u64, u64, i32]
        L.okd_build_tree_canonical.restype = u64
        L.okd_build_tree_canonical.argtypes = [vp, u64, vp, vp, u64, u64, i32, i32]
        L.okd_calc_accel.argtypes = [u64, vp, vp, f64, vp]
        L.okd_calc_accel_all.argtypes = [u64, vp, vp, f64, vp, vp, i32]
        L.okd_kick_drift.argtypes = [u64, vp, vp, f64, i32]
        L.okd_simple_sim.restype = i32
        L.okd_simple_sim.argtypes = [vp, u64, f64, i64, u64, f64, i32, i32, u64, i32]
        L.okd_check_tree_struct.restype = u64
        L.okd_check_tree_struct.argtypes = [vp, vp, i32]
        L.okd_print_tree.restype = i32
        L.okd_print_tree.argtypes = [C.c_char_p, vp, u64, vp]
        L.okd_max_threads.restype = i32
        L.okd_sizeof_node.restype = u64
        assert L.okd_sizeof_node() == self.NODE.itemsize, (L.okd_sizeof_node(), self.NODE.itemsize)

    # -- particles
    def two_bodies(self) -> np.ndarray:
        out = np.zeros(2, PARTICLE)
        self.lib.okd_two_bodies(out.ctypes.data)
        return out

    def circular_orbits(self, n: int, seed: int = 12345) -> np.ndarray:
        out = np.zeros(n + 1, PARTICLE)
        self.lib.okd_circular_orbits(n, seed, out.ctypes.data)
        return out

    def max_threads(self) -> int:
        return int(self.lib.okd_max_threads())

    # -- allocation
    def nodes_needed_for_particles(self, n: int, max_parts: int = 8) -> int:
        return int(self.lib.okd_nodes_needed_for_particles(n, max_parts))

    def allocate_node_vec(self, count: int) -> np.ndarray:
        nodes = np.zeros(count, self.NODE)
        self.lib.okd_fill_default_nodes(nodes.ctypes.data, count)
        return nodes

    # -- selection
    def quickstat_index(self, indices: np.ndarray, goal: int, vals: np.ndarray, seed: int = 1) -> None:
        assert indices.dtype == np.uint64 and vals.dtype == np.float64
        st = C.c_uint64(seed)
        self.lib.okd_quickstat_index_f64(indices.ctypes.data, len(indices), goal, vals.ctypes.data, C.byref(st))

    # -- builds
    def build_tree(self, parts: np.ndarray, max_parts: int = 8, seed: int = 1, cap: int | None = None):
        """array_kd_tree.rs:63-130 (dense, faithful). Returns (nodes, indices, last_used)."""
        n = len(parts)
        cap = cap if cap is not None else 2 * (n // max(1, max_parts // 2) + 1) + 2
        nodes = self.allocate_node_vec(cap)
        idx = np.arange(n, dtype=np.uint64)
        st = C.c_uint64(seed)
        last = self.lib.okd_build_tree(idx.ctypes.data, 0, n, parts.ctypes.data, 0, nodes.ctypes.data, cap, max_parts, C.byref(st))
        assert last != int(U64MAX), "node capacity exceeded"
        return nodes, idx, int(last)

    def build_tree_par4(self, parts: np.ndarray, max_parts: int = 8, seed: int = 1, threads: int = 1):
        """array_kd_tree.rs:515-583 (padded, faithful). Returns (nodes, indices)."""
        n = len(parts)
        nodes = self.allocate_node_vec(self.nodes_needed_for_particles(n, max_parts))
        idx = np.arange(n, dtype=np.uint64)
        self.lib.okd_build_tree_par4(idx.ctypes.data, n, 0, parts.ctypes.data, nodes.ctypes.data, max_parts, seed, threads)
        return nodes, idx

    def build_tree_canonical(self, parts: np.ndarray, max_parts: int = 8, layout: int = LAYOUT_PADDED, threads: int = 1):
        n = len(parts)
        if layout == LAYOUT_PADDED:
            cap = self.nodes_needed_for_particles(n, max_parts)
        else:
            cap = 2 * (n // max(1, max_parts // 2) + 1) + 2
        nodes = self.allocate_node_vec(cap)
        idx = np.arange(n, dtype=np.uint64)
        last = self.lib.okd_build_tree_canonical(idx.ctypes.data, n, parts.ctypes.data, nodes.ctypes.data, cap, max_parts, layout, threads)
        assert last != int(U64MAX), "node capacity exceeded"
        return nodes, idx, int(last)

    # -- walk / kick / sim
    def calc_accel_all(self, parts: np.ndarray, nodes: np.ndarray, theta: float = 0.3, counts: bool = False, threads: int = 0):
        n = len(parts)
        acc = np.zeros((n, 3), np.float64)
        cnt = np.zeros(n, WALK_COUNTS) if counts else None
        threads = threads or self.max_threads()
        self.lib.okd_calc_accel_all(n, parts.ctypes.data, nodes.ctypes.data, theta, acc.ctypes.data,
                                    cnt.ctypes.data if counts else None, threads)
        return (acc, cnt) if counts else acc

    def kick_drift(self, parts: np.ndarray, acc: np.ndarray, dt: float, threads: int = 1) -> None:
        assert acc.dtype == np.float64 and acc.flags.c_contiguous
        self.lib.okd_kick_drift(len(parts), parts.ctypes.data, acc.ctypes.data, dt, threads)

    def simple_sim(self, parts: np.ndarray, dt: float, steps: int, max_parts: int = 8, theta: float = 0.3,
                   layout: int = LAYOUT_PADDED, order: int = ORDER_FAITHFUL, seed: int = 1, threads: int = 0) -> None:
        threads = threads or self.max_threads()
        rc = self.lib.okd_simple_sim(parts.ctypes.data, len(parts), dt, steps, max_parts, theta, layout, order, seed, threads)
        assert rc == 0, f"okd_simple_sim rc={rc}"

    def check_tree_struct(self, nodes: np.ndarray, parts: np.ndarray, dims: int = 3) -> int:
        return int(self.lib.okd_check_tree_struct(nodes.ctypes.data, parts.ctypes.data, dims))

    def print_tree(self, path: str, nodes: np.ndarray, parts: np.ndarray) -> None:
        rc = self.lib.okd_print_tree(path.encode(), nodes.ctypes.data, len(nodes), parts.ctypes.data)
        assert rc == 0
