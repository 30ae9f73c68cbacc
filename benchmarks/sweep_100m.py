#!/usr/bin/env python
"""BASELINE configs[4]: N = 100,000,000 ring, THETA x MAX_PARTS sweep on the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 \
        benchmarks/sweep_100m.py [--number 100000000] [--steps 2] [--thetas 0.3,0.5] [--max-parts 8,16]

The reference has THETA and MAX_PARTS as compile-time constants (Parallel/RustVersion/src/array_kd_tree.rs:14-15); here
they are context parameters.  Every rank generates only its own slice of the initial conditions (host memory: 64 B x
N / ranks) and uploads it with the sharded call; the device holds the full replicated state (~340 B per particle).
One JSON line per configuration (rank 0): particle-steps/s, stage times of the profiled step graph, counted walk work."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multilanguagekdtree_b200 as kd  # noqa: E402
from multilanguagekdtree_b200.array_particle import _splitmix64_stream  # noqa: E402


def ring_slice(n: int, seed: int, first: int, count: int) -> np.ndarray:
    """Records [first, first + count) of circular_orbits(n, seed) without generating the rest."""
    out = np.zeros(count, kd.PARTICLE)
    lo, hi = max(first, 1), first + count          # ring bodies are records 1 .. n
    if first == 0 and count > 0:
        out[0]["r"], out[0]["m"] = 0.00465047, 1.0
    if hi > lo:
        i = np.arange(lo - 1, hi - 1, dtype=np.float64)
        d = 0.1 + (i * 5.0 / float(n))
        v = np.sqrt(1.0 / d)
        # element k of the stream belongs to ring body k - 1: generate [lo .. hi) of it
        with np.errstate(over="ignore"):
            k = np.arange(lo, hi, dtype=np.uint64)
            z = np.uint64(seed) + k * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        theta = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53 * 6.28
        ring = out[lo - first:]
        ring["p"][:, 0] = d * np.cos(theta)
        ring["p"][:, 1] = d * np.sin(theta)
        ring["v"][:, 0] = -v * np.sin(theta)
        ring["v"][:, 1] = v * np.cos(theta)
        ring["m"] = 1e-14
        ring["r"] = 1e-7
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--number", type=int, default=100_000_000)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--thetas", default="0.3,0.5")
    ap.add_argument("--max-parts", default="8,16")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.number
    assert np.array_equal(ring_slice(1000, 12345, 0, 1001), kd.circular_orbits(1000, seed=12345))   # the slice generator is the generator
    first, cnt = kd.host_shard_range(n + 1, rank, world)
    shard = ring_slice(n, 12345, first, cnt)
    for theta in [float(x) for x in a.thetas.split(",")]:
        for mp in [int(x) for x in a.max_parts.split(",")]:
            sim = kd.KDTreeSim(max_parts=mp, theta=theta, device=local, flags=kd.FLAG_PROFILE)
            if world > 1:
                ids = [kd.KDTreeSim.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                sim.comm_init(ids[0], rank, world)
                sim.upload_sharded(shard, n + 1)
            else:
                sim.upload(shard)
            sim.simple_sim(1e-3, 3)            # plain step + capture + one replay
            sim.stage_reset()
            sim.simple_sim(1e-3, a.steps)
            st, k = sim.stage_ms()
            sim.close()
            per = {s: st[s] / k for s in st}
            ms = sum(per.values())
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(json.dumps({"n": n + 1, "gpus": world, "theta": theta, "max_parts": mp, "steps": k,
                                  "ms_per_step": float(t.item()), "particle_steps_per_sec": (n + 1) / (float(t.item()) * 1e-3),
                                  "stage_ms_per_step": per,
                                  "timer": "sum of the stage times of the profiled step graph (CUDA events), max over ranks"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
