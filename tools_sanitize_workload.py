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
print("sanitizer workload done")
