"""World-size-2 `gloo` test (CPU) of the host-side logic of the multi-GPU path: the shard ranges every rank derives
from kdnb_shard_range partition the tree slots, are warp-group aligned and equal-sized (what ncclAllGather needs),
and an all_gather of per-rank slices of a tree-ordered array reproduces the full array — the exchange the library
performs on the GPU with ncclAllGather (kdnb_api.cu: exchange())."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import multilanguagekdtree_b200 as kd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = kd.shard_range(n, rank, world)
        ranges = [None] * world
        dist.all_gather_object(ranges, (b, e))
        shard = ((n + world - 1) // world + 63) // 64 * 64   # kdnb_api.cu: shard_slots_for()
        assert shard % 64 == 0
        for r, (rb, re_) in enumerate(ranges):
            assert rb == min(n, r * shard) and re_ == min(n, (r + 1) * shard)
        assert ranges[-1][1] == n and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        # the exchange: every rank owns acc_t[b:e] (tree order, 3 doubles per slot), padded to equal shards
        full = torch.arange(3 * n, dtype=torch.float64).reshape(n, 3)
        mine = torch.zeros(shard, 3, dtype=torch.float64)
        mine[: e - b] = full[b:e] * 1.0
        gathered = [torch.zeros(shard, 3, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, mine)
        got = torch.cat(gathered)[:n]
        assert torch.equal(got, full)
        # the sharded tree build: every rank's subtree plan (slots and node indices), gathered, tiles the tree
        plans = [None] * world
        dist.all_gather_object(plans, kd.build_shard_plan(n, rank, world))
        assert plans[0][0] == 0 and sum(p[1] for p in plans) == n
        assert all(plans[i][0] + plans[i][1] == plans[i + 1][0] for i in range(world - 1))
        assert all(plans[i][2] + plans[i][3] <= plans[i + 1][2] for i in range(world - 1))     # node ranges disjoint, in order
        assert plans[-1][2] + plans[-1][3] <= kd.nodes_needed_for_particles(n) or n <= 8
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1_000_001, 12345, 63])
def test_shard_ranges_and_exchange_world2(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"


def test_shard_range_pure():
    for n in (1, 63, 64, 65, 1000, 1_000_001, 10_000_001):
        for world in (1, 2, 4, 8):
            covered = 0
            for r in range(world):
                b, e = kd.shard_range(n, r, world)
                assert b == covered or b == n
                covered = max(covered, e)
            assert covered == n
    with pytest.raises(kd.KdnbError):
        kd.shard_range(10, 2, 2)


def test_build_shard_plan_matches_the_oracle_tree():
    """The closed forms behind the sharded build (kdnb_build_shard_plan) against an actual tree of the CPU oracle: the
    level-k segments' slot ranges and node indices, padded and dense layouts, several sizes and world sizes."""
    from oracle.okd import LAYOUT_DENSE as O_DENSE, LAYOUT_PADDED as O_PADDED, Oracle
    orc = Oracle()
    for n in (4097, 20001, 33333):
        parts = orc.circular_orbits(n - 1, seed=n)
        for layout, olayout in ((kd.LAYOUT_PADDED, O_PADDED), (kd.LAYOUT_DENSE, O_DENSE)):
            nodes, idx, _ = orc.build_tree_canonical(parts, layout=olayout)
            for world in (2, 4, 8):
                k = world.bit_length() - 1
                # walk the oracle's tree down k levels: (node, first slot, length) of every level-k segment
                segs = [(0, 0, n)]
                for _ in range(k):
                    nxt = []
                    for node, a, ln in segs:
                        assert nodes["is_internal"][node]
                        left, right = int(nodes["left"][node]), int(nodes["right"][node])
                        nxt += [(left, a, ln // 2), (right, a + ln // 2, ln - ln // 2)]
                    segs = nxt
                for r, (node, a, ln) in enumerate(segs):
                    first_slot, slots, first_node, cnt = kd.build_shard_plan(n, r, world, layout=layout)
                    assert (first_slot, slots, first_node) == (a, ln, node), (n, layout, world, r)
                    nxt_node = segs[r + 1][0] if r + 1 < len(segs) else None
                    if nxt_node is not None:
                        assert first_node + cnt <= nxt_node
    with pytest.raises(kd.KdnbError):
        kd.build_shard_plan(1000, 0, 3)
