#!/usr/bin/env python
"""Full-size property check at N=10,000,000 (BASELINE configs[3] size), run by hand under gpurun (too slow for pytest on
the CPU side: the oracle needs ~1 min per build+walk at this size):  python tests/bigcheck.py [N]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multilanguagekdtree_b200 as kd  # noqa: E402
from oracle.okd import Oracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
orc = Oracle()
parts = orc.circular_orbits(n, seed=12345)
with kd.KDTreeSim(flags=kd.FLAG_WALK_COUNTS) as sim:
    sim.upload(parts)
    t0 = time.time(); sim.build_tree(); sim.synchronize(); print("gpu build s", time.time() - t0)
    nodes, idx = sim.tree()
    t0 = time.time(); sim.calc_accel(); sim.synchronize(); print("gpu walk s", time.time() - t0)
    acc = sim.accel()
    cnt = sim.walk_counts()
with kd.KDTreeSim() as sim:                      # the production walk variant (no counters, planar shortcut)
    sim.upload(parts)
    sim.build_tree()
    sim.calc_accel()
    acc_fast = sim.accel()
assert np.array_equal(acc_fast, acc), "production walk kernel differs from the counting variant"
print("production walk == counting walk, bit for bit")
assert np.array_equal(np.sort(idx), np.arange(n + 1, dtype=np.uint64))
assert len(nodes) == orc.nodes_needed_for_particles(n + 1, 8)
t0 = time.time(); cn, cidx, _ = orc.build_tree_canonical(parts, threads=orc.max_threads()); print("oracle build s", time.time() - t0)
internal = cn["is_internal"].astype(bool)
assert np.array_equal(nodes["kind"] == kd.INTERNAL, internal)
assert np.array_equal(idx, cidx)
for f in ("split_val", "size", "m", "cm"):
    assert np.array_equal(nodes[f][internal], cn[f][internal]), f
for f in ("split_dim", "left", "right"):
    assert np.array_equal(nodes[f][internal].astype(np.uint64), cn[f][internal]), f
print("tree bit-exact at n =", n + 1)
t0 = time.time(); oacc, ocnt = orc.calc_accel_all(parts, cn, counts=True); print("oracle walk s", time.time() - t0)
rel = np.linalg.norm(acc - oacc, axis=1) / np.linalg.norm(oacc, axis=1)
print("acc max rel err", rel.max())
assert rel.max() <= 1e-12
for k, f in enumerate(("node_visits", "accepts", "leaf_visits", "pp")):
    assert np.array_equal(cnt[:, k], ocnt[f]), f
print("walk decisions exact; mean counts", cnt.mean(axis=0))
print("bigcheck ok")
