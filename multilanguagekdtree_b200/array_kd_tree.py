"""Host-side mirror of Parallel/RustVersion/src/array_kd_tree.rs over the C ABI of libkdnb.so.

Same names and argument meaning as the reference's `pub` items; the work runs on the GPU through
include/kdnb.h.  There is no CPU fallback: without the built CUDA library or without a device every call raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (FLAG_EXACT_MATH, FLAG_PROFILE, FLAG_WALK_COUNTS, INTERNAL, LAYOUT_DENSE, LAYOUT_PADDED, LEAF, NODE,
                   NO_INDEX, PARTICLE, STAGES, Config)

MAX_PARTS = 8  # array_kd_tree.rs:14
THETA = 0.3    # array_kd_tree.rs:15


class KdnbError(RuntimeError):
    pass


def nodes_needed_for_particles(num_parts: int, max_parts: int = MAX_PARTS) -> int:
    """array_kd_tree.rs:45-53"""
    return int(_lib.load().kdnb_nodes_needed(num_parts, max_parts))


def allocate_node_vec(num_parts: int, max_parts: int = MAX_PARTS) -> np.ndarray:
    """array_kd_tree.rs:55-60 — every slot is the default Leaf{0, NEGS}."""
    nodes = np.zeros(nodes_needed_for_particles(num_parts, max_parts), NODE)
    nodes["kind"] = LEAF
    nodes["leaf_first"] = NO_INDEX
    return nodes


def shard_range(count: int, rank: int, world: int):
    """Tree-slot range [begin, end) that rank `rank` of `world` walks (kdnb_shard_range; pure, needs no GPU)."""
    b, e = C.c_uint64(0), C.c_uint64(0)
    rc = _lib.load().kdnb_shard_range(count, rank, world, C.byref(b), C.byref(e))
    if rc != 0:
        raise KdnbError(f"kdnb_shard_range rc={rc}")
    return int(b.value), int(e.value)


def build_shard_plan(count: int, rank: int, world: int, max_parts: int = MAX_PARTS, layout: int = LAYOUT_PADDED):
    """What rank `rank` builds in a sharded tree build: (first_slot, slots, first_node, nodes) of its subtree below the top
    log2(world) levels (kdnb_build_shard_plan; pure, needs no GPU)."""
    v = [C.c_uint64(0) for _ in range(4)]
    rc = _lib.load().kdnb_build_shard_plan(count, max_parts, layout, rank, world, *[C.byref(x) for x in v])
    if rc != 0:
        raise KdnbError(f"kdnb_build_shard_plan rc={rc}")
    return tuple(int(x.value) for x in v)


def host_shard_range(total: int, rank: int, world: int):
    """Slice [first, first+count) of the global particle array that rank `rank` keeps on the host (kdnb_host_shard_range)."""
    f, n = C.c_uint64(0), C.c_uint64(0)
    rc = _lib.load().kdnb_host_shard_range(total, rank, world, C.byref(f), C.byref(n))
    if rc != 0:
        raise KdnbError(f"kdnb_host_shard_range rc={rc}")
    return int(f.value), int(n.value)


class KDTreeSim:
    """One GPU context: owns what `simple_sim` owns (acc, tree, indices; array_kd_tree.rs:624-630)."""

    def __init__(self, max_parts: int = MAX_PARTS, theta: float = THETA, layout: int = LAYOUT_PADDED, device: int = 0,
                 flags: int = 0):
        self._L = _lib.load()
        cfg = Config(C.sizeof(Config), device, max_parts, layout, theta, flags, 0)
        self._h = self._L.kdnb_create(C.byref(cfg))
        if not self._h:
            raise KdnbError("kdnb_create: " + self._L.kdnb_last_error(None).decode())
        self.max_parts, self.theta, self.layout, self.flags = max_parts, theta, layout, flags

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.kdnb_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc: int, what: str) -> None:
        if rc != 0:
            raise KdnbError(f"{what}: rc={rc}: {self._L.kdnb_last_error(self._h).decode()}")

    # ---- state
    def upload(self, bodies: np.ndarray) -> None:
        assert bodies.dtype == PARTICLE and bodies.flags.c_contiguous
        self._ck(self._L.kdnb_upload_particles(self._h, bodies.ctypes.data, len(bodies)), "kdnb_upload_particles")

    def download(self, out: np.ndarray | None = None) -> np.ndarray:
        n = self.count
        if out is None:
            out = np.zeros(n, PARTICLE)
        assert out.dtype == PARTICLE and out.flags.c_contiguous
        self._ck(self._L.kdnb_download_particles(self._h, out.ctypes.data, len(out)), "kdnb_download_particles")
        return out

    @property
    def count(self) -> int:
        return int(self._L.kdnb_particle_count(self._h))

    @property
    def node_count(self) -> int:
        return int(self._L.kdnb_node_count(self._h))

    # ---- stages
    def build_tree(self) -> None:
        self._ck(self._L.kdnb_build_tree(self._h), "kdnb_build_tree")

    def calc_accel(self) -> None:
        self._ck(self._L.kdnb_calc_accel(self._h), "kdnb_calc_accel")

    def kick_drift(self, dt: float) -> None:
        self._ck(self._L.kdnb_kick_drift(self._h, dt), "kdnb_kick_drift")

    def simple_sim(self, dt: float, steps: int) -> None:
        self._ck(self._L.kdnb_simple_sim(self._h, dt, steps), "kdnb_simple_sim")

    def simple_sim_bodies(self, bodies: np.ndarray, dt: float, steps: int) -> None:
        assert bodies.dtype == PARTICLE and bodies.flags.c_contiguous
        self._ck(self._L.kdnb_simple_sim_bodies(self._h, bodies.ctypes.data, len(bodies), dt, steps), "kdnb_simple_sim_bodies")

    def simple_sim_bodies_sharded(self, shard: np.ndarray, total: int, dt: float, steps: int) -> None:
        """Multi-GPU form: `shard` is this rank's slice host_shard_range(total, rank, world) of the global bodies."""
        assert shard.dtype == PARTICLE and shard.flags.c_contiguous
        self._ck(self._L.kdnb_simple_sim_bodies_sharded(self._h, shard.ctypes.data, total, dt, steps), "kdnb_simple_sim_bodies_sharded")

    def upload_sharded(self, shard: np.ndarray, total: int) -> None:
        """Multi-GPU: upload this rank's host slice; the other slices arrive over NVLink (kdnb_upload_particles_sharded)."""
        assert shard.dtype == PARTICLE and shard.flags.c_contiguous
        self._ck(self._L.kdnb_upload_particles_sharded(self._h, shard.ctypes.data, total), "kdnb_upload_particles_sharded")

    def synchronize(self) -> None:
        self._ck(self._L.kdnb_synchronize(self._h), "kdnb_synchronize")

    # ---- results
    def accel(self) -> np.ndarray:
        acc = np.zeros((self.count, 3), np.float64)
        self._ck(self._L.kdnb_download_accel(self._h, acc.ctypes.data), "kdnb_download_accel")
        return acc

    def set_accel(self, acc: np.ndarray) -> None:
        acc = np.ascontiguousarray(acc, np.float64)
        assert acc.shape == (self.count, 3)
        self._ck(self._L.kdnb_upload_accel(self._h, acc.ctypes.data), "kdnb_upload_accel")

    def tree(self):
        """(nodes, indices): `Vec<KDTree>` in the reference's index layout and the permuted `indices`."""
        nodes = np.zeros(self.node_count, NODE)
        idx = np.zeros(self.count, np.uint64)
        nn = C.c_uint64(0)
        self._ck(self._L.kdnb_download_tree(self._h, nodes.ctypes.data, len(nodes), C.byref(nn), idx.ctypes.data), "kdnb_download_tree")
        assert nn.value == len(nodes)
        return nodes, idx

    def walk_counts(self) -> np.ndarray:
        cnt = np.zeros((self.count, 4), np.uint64)
        self._ck(self._L.kdnb_download_walk_counts(self._h, cnt.ctypes.data), "kdnb_download_walk_counts")
        return cnt

    # ---- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = _lib.load().kdnb_comm_unique_id(buf)
        if rc != 0:
            raise KdnbError(f"kdnb_comm_unique_id rc={rc}")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int) -> None:
        self._ck(self._L.kdnb_comm_init(self._h, unique_id, rank, world), "kdnb_comm_init")

    # ---- measurement
    def stage_ms(self):
        ms = (C.c_double * 4)()
        steps = C.c_uint64(0)
        self._ck(self._L.kdnb_stage_ms(self._h, ms, C.byref(steps)), "kdnb_stage_ms")
        return dict(zip(STAGES, list(ms))), int(steps.value)

    def stage_reset(self) -> None:
        self._ck(self._L.kdnb_stage_reset(self._h), "kdnb_stage_reset")

    @property
    def launch_count(self) -> int:
        return int(self._L.kdnb_launch_count(self._h))

    def fp64_peak_tflops(self) -> float:
        v = C.c_double(0)
        self._ck(self._L.kdnb_measure_fp64_peak(self._h, C.byref(v)), "kdnb_measure_fp64_peak")
        return v.value

    def flush_l2(self) -> None:
        self._ck(self._L.kdnb_flush_l2(self._h), "kdnb_flush_l2")

    def stopwatch_begin(self) -> None:
        self._ck(self._L.kdnb_device_ms(self._h, 0, None), "kdnb_device_ms")

    def stopwatch_end(self) -> float:
        v = C.c_double(0)
        self._ck(self._L.kdnb_device_ms(self._h, 1, C.byref(v)), "kdnb_device_ms")
        return v.value


def leaf_parts(nodes: np.ndarray, indices: np.ndarray, max_parts: int = MAX_PARTS) -> np.ndarray:
    """Materialise `leaf_parts: [usize; MAX_PARTS]` of every node as the reference stores it:
    the leaf's indices then 0 padding (array_kd_tree.rs:525-529); usize::MAX for slots never written (:16, :58)."""
    out = np.zeros((len(nodes), max_parts), np.uint64)
    unused = (nodes["kind"] == LEAF) & (nodes["leaf_first"] == NO_INDEX)
    out[unused] = NO_INDEX
    for i in np.nonzero((nodes["kind"] == LEAF) & ~unused)[0]:
        f, k = int(nodes["leaf_first"][i]), int(nodes["num_parts"][i])
        out[i, :k] = indices[f:f + k]
    return out


def print_tree(step: int, tree: np.ndarray, indices: np.ndarray, particles: np.ndarray, directory: str = ".") -> str:
    """array_kd_tree.rs:666-692 — write `tree{step}.txt` in the reference's text format (the input of
    TreeVisualizer/src/main/scala/ViewTrees.scala:32-80): the node count, then per node either `L n` followed by n
    lines `x y z`, or `I split_dim split_val left right`.  Numbers are printed like Rust's `{}` (shortest decimal
    that round-trips, never an exponent)."""
    import os

    def f(x: float) -> str:
        return np.format_float_positional(float(x), unique=True, trim="-")

    path = os.path.join(directory, f"tree{step}.txt")
    with open(path, "w") as out:
        out.write(f"{len(tree)}\n")
        for nd in tree:
            if nd["kind"] == INTERNAL:
                out.write(f"I {int(nd['split_dim'])} {f(nd['split_val'])} {int(nd['left'])} {int(nd['right'])}\n")
            else:
                k = int(nd["num_parts"])
                out.write(f"L {k}\n")
                first = int(nd["leaf_first"]) if k else 0
                for i in range(k):
                    p = particles[int(indices[first + i])]["p"]
                    out.write(f"{f(p[0])} {f(p[1])} {f(p[2])}\n")
    return path


# ---- the reference's free functions -------------------------------------------------------------------------

def build_tree_par4(indices: np.ndarray, cur_node: int, particles: np.ndarray, nodes: np.ndarray, thread_cnt: int = 1,
                    max_parts: int = MAX_PARTS) -> None:
    """array_kd_tree.rs:515-583.  Builds the whole tree (cur_node must be 0, `indices` must cover all particles);
    `indices` (uint64) and `nodes` (NODE records from allocate_node_vec) are overwritten in place."""
    if cur_node != 0 or len(indices) != len(particles):
        raise KdnbError("the GPU build constructs the whole tree: cur_node must be 0 and indices must cover all particles")
    with KDTreeSim(max_parts=max_parts, layout=LAYOUT_PADDED) as sim:
        sim.upload(particles)
        sim.build_tree()
        t, idx = sim.tree()
    if len(nodes) < len(t):
        raise KdnbError("nodes is shorter than allocate_node_vec(len(particles))")
    nodes[: len(t)] = t
    indices[:] = idx


def build_tree(indices: np.ndarray, start: int, end: int, particles: np.ndarray, cur_node: int, nodes: np.ndarray,
               max_parts: int = MAX_PARTS):
    """array_kd_tree.rs:63-130 (dense layout).  Returns (last node index used, nodes) — the reference grows `nodes`
    on demand (:75, :123), so the possibly re-allocated array is returned."""
    if cur_node != 0 or start != 0 or end != len(particles):
        raise KdnbError("the GPU build constructs the whole tree: start=0, end=len(particles), cur_node=0")
    with KDTreeSim(max_parts=max_parts, layout=LAYOUT_DENSE) as sim:
        sim.upload(particles)
        sim.build_tree()
        t, idx = sim.tree()
    if len(nodes) < len(t):
        grown = np.zeros(len(t), NODE)
        grown["leaf_first"] = NO_INDEX
        nodes = grown
    nodes[: len(t)] = t
    indices[:] = idx
    return len(t) - 1, nodes


def calc_accel_all(particles: np.ndarray, max_parts: int = MAX_PARTS, theta: float = THETA) -> np.ndarray:
    """acc[i] = calc_accel(i, particles, tree) for every i (array_kd_tree.rs:647 with :585-621) on a fresh tree."""
    with KDTreeSim(max_parts=max_parts, theta=theta) as sim:
        sim.upload(particles)
        sim.build_tree()
        sim.calc_accel()
        return sim.accel()


def simple_sim(bodies: np.ndarray, dt: float, steps: int, max_parts: int = MAX_PARTS, theta: float = THETA) -> None:
    """array_kd_tree.rs:623-664 — advances `bodies` in place by `steps` steps of size `dt`."""
    with KDTreeSim(max_parts=max_parts, theta=theta) as sim:
        sim.simple_sim_bodies(bodies, dt, steps)
