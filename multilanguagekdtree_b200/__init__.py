"""multilanguagekdtree_b200 — the kD-tree N-body step of MarkCLewis/MultiLanguageKDTree (Parallel/RustVersion)
on NVIDIA B200: hand-written sm_100a CUDA (csrc/) behind the C ABI of include/kdnb.h, plus this thin host-side
mirror of the reference's Rust modules (array_particle, array_kd_tree; simd_particle, simd_kd_tree of the Sequential
crate).  No CPU fallback."""
from . import array_kd_tree, array_particle, quickstat, simd_kd_tree, simd_particle
from ._lib import (FLAG_EXACT_MATH, FLAG_PROFILE, FLAG_WALK_COUNTS, INTERNAL, LAYOUT_DENSE, LAYOUT_PADDED, LEAF, NODE,
                   NO_INDEX, PARTICLE)
from .array_kd_tree import (MAX_PARTS, THETA, KDTreeSim, KdnbError, allocate_node_vec, build_shard_plan, build_tree, build_tree_par4,
                            host_shard_range,
                            calc_accel_all, leaf_parts, nodes_needed_for_particles, print_tree, shard_range,
                            simple_sim)
from .array_particle import circular_orbits, two_bodies
from .quickstat import quickstat_index

__all__ = [
    "array_kd_tree", "array_particle", "simd_kd_tree", "simd_particle", "KDTreeSim", "KdnbError", "MAX_PARTS", "THETA", "PARTICLE", "NODE", "LEAF",
    "INTERNAL", "NO_INDEX", "LAYOUT_PADDED", "LAYOUT_DENSE", "FLAG_PROFILE", "FLAG_WALK_COUNTS", "FLAG_EXACT_MATH",
    "allocate_node_vec", "nodes_needed_for_particles", "build_tree", "build_tree_par4", "calc_accel_all", "simple_sim",
    "leaf_parts", "host_shard_range", "build_shard_plan", "print_tree", "shard_range", "circular_orbits", "two_bodies", "quickstat",
    "quickstat_index",
]
