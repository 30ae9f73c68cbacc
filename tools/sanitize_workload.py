import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import multilanguagekdtree_b200 as kd
for n, mp, layout in ((5000, 8, kd.LAYOUT_PADDED), (3000, 7, kd.LAYOUT_DENSE), (9, 8, kd.LAYOUT_PADDED), (70000, 16, kd.LAYOUT_PADDED)):
    parts = kd.circular_orbits(n, seed=n)
    parts["p"][:, 2] = np.random.default_rng(1).normal(size=n + 1) * 0.01 if n == 3000 else 0.0
    with kd.KDTreeSim(max_parts=mp, layout=layout, flags=kd.FLAG_WALK_COUNTS) as sim:
        sim.upload(parts); sim.build_tree(); sim.calc_accel(); sim.walk_counts(); sim.tree(); sim.kick_drift(1e-3)
        sim.simple_sim(1e-3, 4); out = sim.download()
    with kd.KDTreeSim(max_parts=mp, layout=layout) as sim:
        sim.simple_sim_bodies(parts, 1e-3, 4)
    assert np.isfinite(out["p"]).all()
# inputs beyond the 32-bit sort keys: fix-up of equal-key runs, 64-bit fallback (plain launches and conditional graph node)
rng = np.random.default_rng(7)
for kind in ("cluster", "pairs"):
    parts = kd.circular_orbits(20000, seed=3)
    parts["m"] *= 1e-6
    if kind == "cluster":
        parts["p"][: 10000, :2] = 0.25 + rng.random((10000, 2)) * 1e-8
    else:
        base = np.sort(rng.random(10000) * 2.0 - 1.0)
        parts["p"][1::2, 0][:10000] = np.nextafter(base, 2.0)
        parts["p"][2::2, 0][:10000] = base
    with kd.KDTreeSim() as sim:
        sim.upload(parts); sim.build_tree(); sim.calc_accel(); sim.simple_sim(1e-5, 5); out = sim.download()
    assert np.isfinite(out["p"]).all()
# the Sequential crate's SIMD particle surface and a profiled context (event-record nodes in the step graph)
from multilanguagekdtree_b200 import simd_kd_tree, simd_particle
b = simd_particle.circular_orbits(4000, seed=2)
simd_kd_tree.simple_sim(b, 1e-3, 4)
with kd.KDTreeSim(flags=kd.FLAG_PROFILE) as sim:
    sim.upload(kd.circular_orbits(30000, seed=9)); sim.simple_sim(1e-3, 5); sim.stage_ms()
# stand-alone quickstat_index
vals = rng.random(300000)
idx = np.arange(len(vals), dtype=np.uint64)
kd.quickstat_index(idx, 123456, vals)
assert vals[idx[123456]] == np.sort(vals)[123456]
print("sanitizer workload done")
