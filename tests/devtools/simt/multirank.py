"""Development aid: the multi-GPU path of libkdnb with the "ranks" as THREADS of this process, every rank's kernels
executed under the SIMT model (see run.py), NCCL replaced by fake_nccl.cpp and cudaIpc handles by plain pointers.

    python tests/devtools/simt/multirank.py [world] [n] [steps]          # peer stores from inside the walk kernel
    SIMT_IPC=0 python tests/devtools/simt/multirank.py 2 3000 3          # ncclAllGather exchange (no peer mapping)

Checks what tests/multigpu_check.py checks on real GPUs: replicated build + sharded walk + exchange + kick/drift give
accelerations and a trajectory that are BIT-IDENTICAL on every rank and identical to a single-rank run, and the
sharded host-buffer calls agree with the replicated ones.  Logic only (shard ranges, flag handshake, buffer parity,
collective call order); NVLink, NCCL and timing are of course not modelled.  Never touches the product."""
import ctypes as C
import os
import subprocess
import sys
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

os.environ["KDNB_NO_GRAPH"] = "1"
os.environ.setdefault("SIMT_THREADS", "3")          # OS threads per kernel launch and rank
import build as simt_build  # noqa: E402

lib = simt_build.build()
fake = os.path.join(HERE, "_build", "libnccl.so.2")
src = os.path.join(HERE, "fake_nccl.cpp")
if True:  # always rebuilt from fake_nccl.cpp (a second of g++): a stale or foreign binary of that name must never be loaded
    tmp = fake + ".%d" % os.getpid()  # (renamed into place: several of these processes may start at once, pytest -n)
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-Wl,-soname,libnccl.so.2", "-o", tmp, src, "-lpthread"],
                   check=True)
    os.replace(tmp, fake)
C.CDLL(fake, mode=C.RTLD_GLOBAL)                     # dlopen("libnccl.so.2") inside the library now finds this one
from multilanguagekdtree_b200 import _lib  # noqa: E402

_lib.LIB_PATH = lib
import multilanguagekdtree_b200 as kd  # noqa: E402


def rank_main(rank, world, uid, ics, steps, out, errs):
    try:
        n1 = len(ics)
        sim = kd.KDTreeSim()
        sim.comm_init(uid, rank, world)
        sim.upload(ics)
        sim.build_tree()
        nodes, idx = sim.tree()
        sim.calc_accel()
        acc = sim.accel()
        l0 = sim.launch_count
        sim.simple_sim(1e-3, steps)
        launches = (sim.launch_count - l0) / steps
        res = sim.download()
        first, cnt = kd.host_shard_range(n1, rank, world)
        shard = ics[first:first + cnt].copy()
        sim.simple_sim_bodies_sharded(shard, n1, 1e-3, steps)
        assert shard.tobytes() == res[first:first + cnt].tobytes(), "sharded host path differs from the replicated path"
        sim.close()
        out[rank] = (acc, res, launches, nodes, idx)
    except BaseException as e:  # noqa: BLE001
        errs.append((rank, repr(e)))
        raise


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    ics = kd.circular_orbits(n, seed=4242)
    uid = kd.KDTreeSim.comm_unique_id()
    out, errs = [None] * world, []
    threads = [threading.Thread(target=rank_main, args=(r, world, uid, ics, steps, out, errs), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(600)
    if errs or any(t.is_alive() for t in threads) or any(o is None for o in out):
        print("multirank FAILED:", errs or "a rank did not finish")
        os._exit(1)                                   # (ranks blocked in a collective cannot be joined)
    with kd.KDTreeSim() as one:
        one.upload(ics)
        one.build_tree()
        nodes1, idx1 = one.tree()
        one.calc_accel()
        acc1 = one.accel()
        one.simple_sim(1e-3, steps)
        res1 = one.download()
    for r in range(world):
        assert out[r][4].tobytes() == idx1.tobytes(), f"rank {r}: tree order differs from the single-rank build"
        assert out[r][3].tobytes() == nodes1.tobytes(), f"rank {r}: node records differ from the single-rank build"
        assert out[r][0].tobytes() == acc1.tobytes(), f"rank {r}: accelerations differ from the single-rank walk"
        assert out[r][1].tobytes() == res1.tobytes(), f"rank {r}: trajectory differs from the single-rank run"
    mode = "ncclAllGather exchange" if os.environ.get("SIMT_IPC") == "0" or os.environ.get("KDNB_NO_P2P") else "peer stores"
    print(f"multirank ok: world={world} n={n + 1} steps={steps} ({mode}; {out[0][2]:.1f} kernel launches per step: the peer mode "
          f"adds the flag-wait kernel): every rank and the single-rank run are bit-identical")


if __name__ == "__main__":
    main()
