"""Host-side mirror of Parallel/RustVersion/src/array_particle.rs: the Particle record and the IC generators.

Particles are numpy structured arrays with the reference's field order (p, v, r, m; 64 bytes), which is exactly
the kdnb_particle the C ABI takes — no conversion at the boundary.
"""
from __future__ import annotations

import numpy as np

from ._lib import PARTICLE

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64_stream(seed: int, count: int) -> np.ndarray:
    """count outputs of splitmix64 started at `seed` (vectorised: state_k = seed + (k+1)*golden)."""
    with np.errstate(over="ignore"):
        k = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + k * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def two_bodies() -> np.ndarray:
    """array_particle.rs:10-17"""
    out = np.zeros(2, PARTICLE)
    out[0]["r"], out[0]["m"] = 1.0, 1.0
    out[1]["p"] = (1.0, 0.0, 0.0)
    out[1]["v"] = (0.0, 1.0, 0.0)
    out[1]["r"], out[1]["m"] = 1e-4, 1e-20
    return out


def circular_orbits(n: int, seed: int = 12345) -> np.ndarray:
    """array_particle.rs:19-44 — returns n+1 particles: the central body and n ring bodies.

    The reference draws the angles from an unseeded fastrand::f64() (:31); here they are the top 53 bits of a
    splitmix64 stream so that runs are reproducible."""
    out = np.zeros(n + 1, PARTICLE)
    out[0]["r"], out[0]["m"] = 0.00465047, 1.0
    i = np.arange(n, dtype=np.float64)
    d = 0.1 + (i * 5.0 / float(n))
    v = np.sqrt(1.0 / d)
    u = (_splitmix64_stream(seed, n) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    theta = u * 6.28
    ring = out[1:]
    ring["p"][:, 0] = d * np.cos(theta)
    ring["p"][:, 1] = d * np.sin(theta)
    ring["v"][:, 0] = -v * np.sin(theta)
    ring["v"][:, 1] = v * np.cos(theta)
    ring["m"] = 1e-14
    ring["r"] = 1e-7
    return out
