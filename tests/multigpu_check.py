#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py [N] [steps]

Every rank builds the (replicated) tree, walks its shard of tree slots, all-gathers the accelerations with NCCL and
kick/drifts everything; the result must be BIT-IDENTICAL on every rank and identical to a single-GPU run, because the
same arithmetic is applied to every particle wherever it is walked.  KDNB_SHARD_BUILD=1 forces the sharded tree build
(every rank builds one subtree below the top log2(world) levels and pushes it to its peers): the tree — node records and
tree order — must then be bit-identical on every rank and to the single-GPU build as well."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multilanguagekdtree_b200 as kd  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ics = kd.circular_orbits(n, seed=4242)
    sim = kd.KDTreeSim(device=local)
    ids = [kd.KDTreeSim.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sim.comm_init(ids[0], rank, world)
    sim.upload(ics)
    sim.build_tree()
    nodes, idx = sim.tree()
    sim.calc_accel()
    acc = sim.accel()
    sim.simple_sim(1e-3, steps)
    out = sim.download()
    # sharded host buffers: upload own shard + NVLink all-gather, run, download own shard
    first, cnt = kd.host_shard_range(n + 1, rank, world)
    shard = ics[first:first + cnt].copy()
    sim.simple_sim_bodies_sharded(shard, n + 1, 1e-3, steps)
    assert shard.tobytes() == out[first:first + cnt].tobytes(), "sharded host path differs from the replicated path"
    sim.close()
    blobs = [None] * world
    import hashlib
    tree_digest = hashlib.sha256(nodes.tobytes() + idx.tobytes()).hexdigest()
    dist.all_gather_object(blobs, (out.tobytes(), acc.tobytes(), tree_digest))
    if rank == 0:
        for r in range(1, world):
            assert blobs[r] == blobs[0], f"rank {r} differs from rank 0"
        with kd.KDTreeSim(device=local) as one:
            one.upload(ics)
            one.build_tree()
            nodes1, idx1 = one.tree()
            one.calc_accel()
            acc1 = one.accel()
            one.simple_sim(1e-3, steps)
            out1 = one.download()
        assert hashlib.sha256(nodes1.tobytes() + idx1.tobytes()).hexdigest() == tree_digest, "tree differs from the single-GPU build"
        assert acc1.tobytes() == acc.tobytes(), "sharded accelerations differ from the single-GPU walk"
        assert out1.tobytes() == out.tobytes(), "sharded trajectory differs from the single-GPU trajectory"
        print(f"multigpu_check ok: world={world} n={n + 1} steps={steps}: all ranks and the 1-GPU run are bit-identical")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
