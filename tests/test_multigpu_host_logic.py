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
