#!/usr/bin/env python
"""Stand-alone order-statistic benchmark — the B200 counterpart of Parallel/RustVersion/src/bin/bench_quickstat.rs
(n = 100,000,000 uniform values, identity indices, one random goal; `bench_quickstat.rs:8-21`).  The reference
publishes, hardware unstated (Parallel/README.md:53-57): sequential quickstat_index 2.250 s, quickstat_index_par 0.536 s.

    python benchmarks/bench_quickstat.py [--number 100000000] [--repeats 3] [--cpu]

Prints one JSON line.  `device_ms` is the selection on the device (radix select + stable three-way partition) with the
arrays resident; `call_s` is the whole C-ABI call with host arrays (pageable upload of vals + indices, download of the
permuted indices).  `--cpu` also times the oracle's restatement of quickstat_index on one host core."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multilanguagekdtree_b200 as kd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--number", "-n", type=int, default=100_000_000)
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    rng = np.random.default_rng(12345)
    vals = rng.random(a.number)
    out = {"bench": "quickstat_index", "n": a.number,
           "reference_published_s": {"quickstat_index": 2.250, "quickstat_index_par": 0.536,
                                     "source": "Parallel/README.md:53-57 (hardware unstated)"}}
    dev, call = [], []
    with kd.KDTreeSim() as sim:
        for r in range(a.repeats + 1):
            idx = np.arange(a.number, dtype=np.uint64)
            goal = int(rng.integers(0, a.number))
            t0 = time.perf_counter()
            ms = kd.quickstat_index(idx, goal, vals, sim=sim)
            t1 = time.perf_counter()
            pivot = vals[idx[goal]]
            assert np.all(vals[idx[:goal]] < pivot) and np.all(vals[idx[goal:]] >= pivot)   # quickstat.rs:199-216
            if r:                                   # the first call pays context warm-up
                dev.append(ms)
                call.append(t1 - t0)
    out["device_ms"] = min(dev)
    out["call_s"] = min(call)
    out["elements_per_s_device"] = a.number / (min(dev) * 1e-3)
    out["algorithmic_bytes_per_element"] = 100   # keys 8+8, eight histogram passes 8 each, partition 2 x (8 + 4)
    out["hbm_gbs_device"] = 100.0 * a.number / (min(dev) * 1e-3) / 1e9
    if a.cpu:
        from oracle.okd import Oracle, build
        build()
        orc = Oracle()
        idx = np.arange(a.number, dtype=np.uint64)
        t0 = time.perf_counter()
        orc.quickstat_index(idx, int(rng.integers(0, a.number)), vals, seed=1)
        out["cpu_port_s"] = {"value": time.perf_counter() - t0, "cores": 1, "kind": "port (oracle restatement of quickstat.rs:9-34)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
