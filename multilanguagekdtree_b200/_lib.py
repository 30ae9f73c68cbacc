"""ctypes binding of libkdnb.so — the declarations of include/kdnb.h, one to one."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# KDNB_LIB: another build of the same library (compile-time A/B variants, tools/ab.sh); default = the in-tree build
LIB_PATH = os.environ.get("KDNB_LIB") or os.path.join(HERE, "libkdnb.so")

# mirrors kdnb_particle == `pub struct Particle` (Parallel/RustVersion/src/array_particle.rs:3-8)
PARTICLE = np.dtype([("p", "<f8", (3,)), ("v", "<f8", (3,)), ("r", "<f8"), ("m", "<f8")], align=True)
# mirrors kdnb_node == flat `pub enum KDTree` (array_kd_tree.rs:18-34)
NODE = np.dtype(
    [("kind", "<u4"), ("split_dim", "<u4"), ("num_parts", "<u8"), ("leaf_first", "<u8"), ("split_val", "<f8"),
     ("m", "<f8"), ("cm", "<f8", (3,)), ("size", "<f8"), ("left", "<u8"), ("right", "<u8")], align=True)
assert PARTICLE.itemsize == 64 and NODE.itemsize == 88

LEAF, INTERNAL = 0, 1
LAYOUT_PADDED, LAYOUT_DENSE = 0, 1
FLAG_PROFILE, FLAG_WALK_COUNTS, FLAG_EXACT_MATH = 1, 2, 4
NO_INDEX = 0xFFFFFFFFFFFFFFFF
STAGES = ("build", "walk", "kick", "exchange")

# every symbol include/kdnb.h declares (tests check the library exports all of them)
SYMBOLS = [
    "kdnb_create", "kdnb_destroy", "kdnb_last_error", "kdnb_version", "kdnb_upload_particles",
    "kdnb_download_particles", "kdnb_particle_count", "kdnb_build_tree", "kdnb_calc_accel", "kdnb_kick_drift",
    "kdnb_simple_sim", "kdnb_simple_sim_host", "kdnb_simple_sim_bodies", "kdnb_synchronize", "kdnb_download_accel",
    "kdnb_upload_accel", "kdnb_download_tree", "kdnb_download_walk_counts", "kdnb_nodes_needed", "kdnb_node_count",
    "kdnb_shard_range", "kdnb_host_shard_range", "kdnb_upload_particles_sharded",
    "kdnb_download_particles_sharded", "kdnb_simple_sim_bodies_sharded", "kdnb_comm_unique_id", "kdnb_comm_init", "kdnb_stage_ms", "kdnb_stage_reset", "kdnb_launch_count",
    "kdnb_measure_fp64_peak", "kdnb_flush_l2", "kdnb_device_ms", "kdnb_host_alloc", "kdnb_host_free",
    "kdnb_quickstat_index", "kdnb_upload_particles_simd", "kdnb_download_particles_simd", "kdnb_simple_sim_bodies_simd", "kdnb_build_shard_plan",
]


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("max_parts", C.c_uint32), ("layout", C.c_int32),
                ("theta", C.c_double), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


_lib = None


def load() -> C.CDLL:
    """Load libkdnb.so.  Fails loudly when the CUDA extension has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m multilanguagekdtree_b200.build` "
                          "(hand-written CUDA for sm_100a; this package has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, i64, f64, i32, u32 = C.c_void_p, C.c_uint64, C.c_int64, C.c_double, C.c_int, C.c_uint32
    L.kdnb_create.restype = vp
    L.kdnb_create.argtypes = [C.POINTER(Config)]
    L.kdnb_destroy.argtypes = [vp]
    L.kdnb_last_error.restype = C.c_char_p
    L.kdnb_last_error.argtypes = [vp]
    L.kdnb_version.restype = i32
    L.kdnb_upload_particles.argtypes = [vp, vp, u64]
    L.kdnb_download_particles.argtypes = [vp, vp, u64]
    L.kdnb_particle_count.restype = u64
    L.kdnb_particle_count.argtypes = [vp]
    L.kdnb_build_tree.argtypes = [vp]
    L.kdnb_calc_accel.argtypes = [vp]
    L.kdnb_kick_drift.argtypes = [vp, f64]
    L.kdnb_simple_sim.argtypes = [vp, f64, i64]
    L.kdnb_simple_sim_host.argtypes = [C.POINTER(Config), vp, u64, f64, i64]
    L.kdnb_simple_sim_bodies.argtypes = [vp, vp, u64, f64, i64]
    L.kdnb_synchronize.argtypes = [vp]
    L.kdnb_download_accel.argtypes = [vp, vp]
    L.kdnb_upload_accel.argtypes = [vp, vp]
    L.kdnb_download_tree.argtypes = [vp, vp, u64, vp, vp]
    L.kdnb_download_walk_counts.argtypes = [vp, vp]
    L.kdnb_nodes_needed.restype = u64
    L.kdnb_nodes_needed.argtypes = [u64, u32]
    L.kdnb_node_count.restype = u64
    L.kdnb_node_count.argtypes = [vp]
    L.kdnb_shard_range.argtypes = [u64, i32, i32, vp, vp]
    L.kdnb_shard_range.restype = i32
    L.kdnb_host_shard_range.argtypes = [u64, i32, i32, vp, vp]
    L.kdnb_host_shard_range.restype = i32
    L.kdnb_upload_particles_sharded.argtypes = [vp, vp, u64]
    L.kdnb_upload_particles_sharded.restype = i32
    L.kdnb_download_particles_sharded.argtypes = [vp, vp]
    L.kdnb_download_particles_sharded.restype = i32
    L.kdnb_simple_sim_bodies_sharded.argtypes = [vp, vp, u64, f64, i64]
    L.kdnb_simple_sim_bodies_sharded.restype = i32
    L.kdnb_comm_unique_id.argtypes = [vp]
    L.kdnb_comm_init.argtypes = [vp, vp, i32, i32]
    L.kdnb_stage_ms.argtypes = [vp, vp, vp]
    L.kdnb_stage_reset.argtypes = [vp]
    L.kdnb_launch_count.restype = u64
    L.kdnb_launch_count.argtypes = [vp]
    L.kdnb_measure_fp64_peak.argtypes = [vp, vp]
    L.kdnb_flush_l2.argtypes = [vp]
    L.kdnb_device_ms.argtypes = [vp, i32, vp]
    L.kdnb_host_alloc.restype = vp
    L.kdnb_host_alloc.argtypes = [u64]
    L.kdnb_host_free.argtypes = [vp]
    L.kdnb_quickstat_index.argtypes = [vp, vp, u64, vp, u64, u64, vp]
    L.kdnb_quickstat_index.restype = i32
    L.kdnb_build_shard_plan.argtypes = [u64, u32, i32, i32, i32, vp, vp, vp, vp]
    L.kdnb_build_shard_plan.restype = i32
    L.kdnb_upload_particles_simd.argtypes = [vp, vp, u64]
    L.kdnb_download_particles_simd.argtypes = [vp, vp, u64]
    L.kdnb_simple_sim_bodies_simd.argtypes = [vp, vp, u64, f64, i64]
    for fn in ("kdnb_upload_particles_simd", "kdnb_download_particles_simd", "kdnb_simple_sim_bodies_simd"):
        getattr(L, fn).restype = i32
    for fn in ("kdnb_upload_particles", "kdnb_download_particles", "kdnb_build_tree", "kdnb_calc_accel", "kdnb_kick_drift",
               "kdnb_simple_sim", "kdnb_simple_sim_host", "kdnb_simple_sim_bodies", "kdnb_synchronize",
               "kdnb_download_accel", "kdnb_upload_accel", "kdnb_download_tree", "kdnb_download_walk_counts",
               "kdnb_shard_range", "kdnb_host_shard_range", "kdnb_upload_particles_sharded",
    "kdnb_download_particles_sharded", "kdnb_simple_sim_bodies_sharded", "kdnb_comm_unique_id", "kdnb_comm_init", "kdnb_stage_ms", "kdnb_stage_reset", "kdnb_measure_fp64_peak",
               "kdnb_flush_l2", "kdnb_device_ms"):
        getattr(L, fn).restype = i32
    _lib = L
    return L
