#!/usr/bin/env python
"""Stand-alone tree-build benchmark — the B200 counterpart of Parallel/RustVersion/src/bin/bench_build_tree.rs
(N = 10,000,000, `bench_build_tree.rs:7`).  The reference publishes, hardware unstated (Parallel/README.md:80-86):
sequential build_tree 11.20 s, build_tree_par4 2.217 s.

    python benchmarks/bench_build_tree.py [--number 10000000] [--repeats 5] [--cpu]

Prints one JSON line.  `--cpu` also times the oracle's restatement of build_tree (sequential) and build_tree_par4
(OpenMP tasks == rayon::join) on the host cores — a reported baseline, not the thing measured."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multilanguagekdtree_b200 as kd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--number", "-n", type=int, default=10_000_000)
    ap.add_argument("--repeats", type=int, default=5)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    parts = kd.circular_orbits(a.number, seed=12345)
    out = {"bench": "build_tree", "n_particles": a.number + 1,
           "reference_published_s": {"build_tree_sequential": 11.20, "build_tree_par4": 2.217,
                                     "source": "Parallel/README.md:80-86 (hardware unstated)"}}
    for name, layout in (("build_tree_par4 (padded layout)", kd.LAYOUT_PADDED), ("build_tree (dense layout)", kd.LAYOUT_DENSE)):
        with kd.KDTreeSim(layout=layout) as sim:
            sim.upload(parts)
            for _ in range(2):
                sim.build_tree()
            sim.synchronize()
            ms = []
            for _ in range(a.repeats):
                sim.stopwatch_begin()
                sim.build_tree()
                ms.append(sim.stopwatch_end())
            out[name] = {"device_ms_min": min(ms), "device_ms_median": sorted(ms)[len(ms) // 2],
                         "particles_per_s": (a.number + 1) / (min(ms) * 1e-3)}
    if a.cpu:
        from oracle.okd import Oracle
        orc = Oracle()
        op = orc.circular_orbits(a.number, seed=12345)
        t0 = time.perf_counter(); orc.build_tree_par4(op, threads=orc.max_threads()); t_par4 = time.perf_counter() - t0
        t0 = time.perf_counter(); orc.build_tree(op); t_seq = time.perf_counter() - t0
        out["cpu_oracle_port_s"] = {"build_tree_sequential": t_seq, "build_tree_par4": t_par4, "threads": orc.max_threads()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
