#!/usr/bin/env python
"""Generate golden vectors from the REFERENCE's own Python implementation of the path.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

It imports /root/reference/Sequential/PythonVersions/PureVersion/src/{kd_tree,particle}.py
UNMODIFIED, replaces two module globals — `kd_tree.MAX_PARTS` (7 there, 8 in the Rust hot path,
Parallel/RustVersion/src/array_kd_tree.rs:14) and `kd_tree.randrange` (the unseeded pivot source,
kd_tree.py:106) with the splitmix64 pivot stream that oracle/kdtree_oracle.c implements — and records
the reference's build_tree / calc_accel / simple_sim outputs.  Initial conditions come from the oracle's
seeded circular_orbits (the reference's are unseeded) and are stored in the fixture.

PureVersion follows the same operation order as the Rust code (sequential node sums, pairwise walk,
kick then drift), so the oracle's dense + faithful mode must reproduce these files BIT FOR BIT
(tests/test_oracle_golden.py).

One case (`pure3d_*`) needs ONE documented source patch: PureVersion never considers z as a split dimension
(`for dim in range(1, 2):  # FIXME: what is this`, kd_tree.py:96), while the Rust hot path loops over all three
(`for dim in 1..3`, Parallel/RustVersion/src/array_kd_tree.rs:552).  For that case the module is loaded from its
source text with exactly that one token changed (`range(1, 2)` -> `range(1, 3)`), nothing else; the initial conditions
get a z extent so that z IS chosen in many nodes.  It pins the 3-D split choice outside the oracle's own restatement.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference/Sequential/PythonVersions/PureVersion/src"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF_SRC)
sys.setrecursionlimit(100000)

import kd_tree as ref_kd  # noqa: E402  (the reference)
import particle as ref_p  # noqa: E402  (the reference)

from oracle.okd import PARTICLE, Oracle  # noqa: E402

MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def randrange(self, s: int, e: int) -> int:
        return s + self.next() % (e - s)


def to_ref(parts: np.ndarray):
    return [ref_p.Particle(tuple(map(float, q["p"])), tuple(map(float, q["v"])), float(q["r"]), float(q["m"])) for q in parts]


def from_ref(bodies) -> np.ndarray:
    out = np.zeros(len(bodies), PARTICLE)
    for i, b in enumerate(bodies):
        out[i]["p"] = (b.p.x, b.p.y, b.p.z)
        out[i]["v"] = (b.v.x, b.v.y, b.v.z)
        out[i]["r"] = b.r
        out[i]["m"] = b.m
    return out


def dump_tree(nodes, last: int, cap: int) -> dict:
    n = last + 1
    d = dict(
        is_internal=np.zeros(n, np.uint8), num_parts=np.zeros(n, np.uint64),
        leaf_parts=np.full((n, cap), 0, np.uint64), split_dim=np.zeros(n, np.uint64),
        split_val=np.zeros(n), m=np.zeros(n), cm=np.zeros((n, 3)), size=np.zeros(n),
        left=np.zeros(n, np.uint64), right=np.zeros(n, np.uint64),
    )
    for i in range(n):
        nd = nodes[i]
        if nd.num_parts > 0:
            d["num_parts"][i] = nd.num_parts
            d["leaf_parts"][i, : nd.num_parts] = nd.particles
        else:
            d["is_internal"][i] = 1
            d["split_dim"][i] = nd.split_dim
            d["split_val"][i] = nd.split_val
            d["m"][i] = nd.m
            d["cm"][i] = (nd.cm.x, nd.cm.y, nd.cm.z)
            d["size"][i] = nd.size
            d["left"][i] = nd.left
            d["right"][i] = nd.right
    return d


def load_3d_variant():
    """kd_tree.py with the split-dimension loop over all three dimensions (the only change, see the module docstring)."""
    import types
    src = open(os.path.join(REF_SRC, "kd_tree.py")).read()
    old = "for dim in range(1, 2):"
    assert src.count(old) == 1, "reference source changed: re-check the patch"
    mod = types.ModuleType("kd_tree_3d")
    mod.__file__ = os.path.join(REF_SRC, "kd_tree.py")
    sys.modules[mod.__name__] = mod  # (dataclasses look their module up by name)
    exec(compile(src.replace(old, "for dim in range(1, 3):"), mod.__file__, "exec"), mod.__dict__)
    return mod


def case(name: str, parts0: np.ndarray, max_parts: int, pivot_seed: int, steps: int, dt: float, ref_kd=ref_kd):
    ref_kd.MAX_PARTS = max_parts
    rng = SplitMix64(pivot_seed)
    ref_kd.randrange = rng.randrange

    bodies = to_ref(parts0)
    sysm = ref_kd.System.from_amount(len(bodies))
    last = sysm.build_tree(0, len(bodies), bodies, 0)
    tree = dump_tree(sysm.nodes, last, max(8, max_parts))
    acc = np.array([[a.x, a.y, a.z] for a in (ref_kd.calc_accel(i, bodies, sysm.nodes) for i in range(len(bodies)))])

    # a fresh pivot stream for the trajectory, as okd_simple_sim(seed) starts one
    rng2 = SplitMix64(pivot_seed)
    ref_kd.randrange = rng2.randrange
    bodies2 = to_ref(parts0)
    ref_kd.simple_sim(bodies2, dt, steps)
    after = from_ref(bodies2)

    out = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(
        out, parts0=parts0, max_parts=max_parts, pivot_seed=pivot_seed, steps=steps, dt=dt,
        indices=np.array(sysm.indices, np.uint64), last=last, acc=acc, after=after,
        **{f"tree_{k}": v for k, v in tree.items()},
    )
    print(f"wrote {out}: n={len(parts0)} last_node={last} steps={steps}")


def main():
    orc = Oracle()
    # the Rust hot path's constants (MAX_PARTS=8) on ring ICs
    case("pure_ring300_mp8", orc.circular_orbits(300, seed=7), 8, 1001, 5, 1e-3)
    # PureVersion's own constant (MAX_PARTS=7 == Sequential/RustVersion, BASELINE config #1 semantics)
    case("pure_ring200_mp7", orc.circular_orbits(200, seed=11), 7, 2002, 3, 1e-3)
    # reference test sizes: two_leaves (12 particles, array_kd_tree.rs:711-731)
    case("pure_ring11_mp8", orc.circular_orbits(11, seed=3), 8, 3003, 4, 1e-3)
    # two_bodies half orbit (Parallel/GoVersion/kdtree_test.go:63-74 prints this; no assertion there)
    case("pure_two_bodies", orc.two_bodies(), 8, 4004, 1000, math.pi / 1000)
    # 3-D: the ring lifted out of its plane (z extent 4 against 10 in x / y, vertical velocities), all three split dimensions
    p3 = orc.circular_orbits(400, seed=19)
    i = np.arange(1, len(p3), dtype=np.float64)
    p3["p"][1:, 2] = 2.0 * np.cos(1.7 * i)
    p3["v"][1:, 2] = 0.05 * np.sin(0.3 * i)
    case("pure3d_ring400_mp8", p3, 8, 5005, 4, 1e-3, ref_kd=load_3d_variant())


if __name__ == "__main__":
    main()
