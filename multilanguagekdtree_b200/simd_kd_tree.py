"""Host-side mirror of Sequential/RustVersion/src/simd_kd_tree.rs over the C ABI: the Sequential crate's step on its
f64x4 particles.  MAX_PARTS = 7 (simd_kd_tree.rs:9), dense `build_tree` layout (:49-138); the arithmetic is the scalar
path's (the fourth lane adds an exact 0 to every sum), so results are bit-identical to array_kd_tree on the same data."""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import LAYOUT_DENSE
from .array_kd_tree import THETA, KDTreeSim
from .simd_particle import PARTICLE_SIMD

MAX_PARTS = 7  # simd_kd_tree.rs:9


def nodes_needed_for_particles(num_parts: int) -> int:
    """simd_kd_tree.rs:41-43"""
    return 2 * (num_parts // (MAX_PARTS // 2) + 1)


def simple_sim(bodies: np.ndarray, dt: float, steps: int, theta: float = THETA) -> None:
    """simd_kd_tree.rs:169-202 — `bodies` (PARTICLE_SIMD) is advanced in place."""
    assert bodies.dtype == PARTICLE_SIMD and bodies.flags.c_contiguous
    with KDTreeSim(max_parts=MAX_PARTS, theta=theta, layout=LAYOUT_DENSE) as sim:
        sim._ck(_lib.load().kdnb_simple_sim_bodies_simd(sim._h, bodies.ctypes.data, len(bodies), dt, steps),
                "kdnb_simple_sim_bodies_simd")
