"""Host-side mirror of Sequential/RustVersion/src/simd_particle.rs: the f64x4 Particle record and its IC generators.

`Particle { p: f64x4, v: f64x4, r: f64, m: f64 }` (simd_particle.rs:3-8) is 96 bytes (f64x4 is 32-byte aligned) — the
kdnb_particle_simd of the C ABI.  Lane 3 of p and v is padding: every generator leaves it 0 and the arithmetic of
simd_kd_tree.rs keeps it 0.  Note the two differences from array_particle.rs that the reference itself has: the ring
angles use TAU instead of 6.28 (simd_particle.rs:40) and two_bodies() uses r = 1.0 / 1e-4 (:13-26).
"""
from __future__ import annotations

import numpy as np

from .array_particle import _splitmix64_stream

PARTICLE_SIMD = np.dtype({"names": ["p", "v", "r", "m"], "formats": [("<f8", (4,)), ("<f8", (4,)), "<f8", "<f8"],
                          "offsets": [0, 32, 64, 72], "itemsize": 96})


def two_bodies() -> np.ndarray:
    """simd_particle.rs:10-27"""
    out = np.zeros(2, PARTICLE_SIMD)
    out[0]["r"], out[0]["m"] = 1.0, 1.0
    out[1]["p"] = (1.0, 0.0, 0.0, 0.0)
    out[1]["v"] = (0.0, 1.0, 0.0, 0.0)
    out[1]["r"], out[1]["m"] = 1e-4, 1e-20
    return out


def circular_orbits(n: int, seed: int = 12345) -> np.ndarray:
    """simd_particle.rs:29-55 — n+1 particles; angles from a seeded splitmix64 stream (the reference's are unseeded)."""
    out = np.zeros(n + 1, PARTICLE_SIMD)
    out[0]["r"], out[0]["m"] = 0.00465047, 1.0
    i = np.arange(n, dtype=np.float64)
    d = 0.1 + (i * 5.0 / float(n))
    v = np.sqrt(1.0 / d)
    u = (_splitmix64_stream(seed, n) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    theta = u * (2.0 * np.pi)  # std::f64::consts::TAU
    ring = out[1:]
    ring["p"][:, 0] = d * np.cos(theta)
    ring["p"][:, 1] = d * np.sin(theta)
    ring["v"][:, 0] = -v * np.sin(theta)
    ring["v"][:, 1] = v * np.cos(theta)
    ring["m"] = 1e-14
    ring["r"] = 1e-7
    return out


def to_scalar(bodies: np.ndarray) -> np.ndarray:
    """The same particles as array_particle records (lanes 0..2)."""
    from ._lib import PARTICLE
    out = np.zeros(len(bodies), PARTICLE)
    out["p"], out["v"], out["r"], out["m"] = bodies["p"][:, :3], bodies["v"][:, :3], bodies["r"], bodies["m"]
    return out
