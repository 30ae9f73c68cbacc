"""Development aid: randomised campaign on top of multirank.py — random input kinds (ring, cube, ties, collinear, flat,
mixed scales, duplicates, lattice, zero masses), world sizes 2-5, sizes 5-6000 (so that some ranks get empty or
ragged shards), 1-3 steps; every rank's accelerations and trajectory must be bit-identical to a single-rank run.

    python tests/devtools/simt/multirank_fuzz.py <first_seed> <count>          (SIMT_IPC=0: ncclAllGather exchange)"""
import os, sys, threading, importlib.util
import numpy as np
sys.argv = [sys.argv[0]] + sys.argv[1:]
first = int(sys.argv[1]); count = int(sys.argv[2])
HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location('mr', os.path.join(HERE, 'multirank.py')); mr = importlib.util.module_from_spec(spec)
sys_argv_saved = sys.argv; sys.argv = ['multirank.py']
spec.loader.exec_module(mr)
sys.argv = sys_argv_saved
kd = mr.kd
from oracle import okd
ORC = okd.Oracle()
KINDS = ["ring", "cube", "quant", "line_x", "flat_y", "two_scales", "dup", "grid", "zero_mass"]
def make(kind, n, rng):
    if kind == "ring":
        return ORC.circular_orbits(max(1, n - 1), seed=int(rng.integers(1, 1 << 30)))
    parts = np.zeros(n, okd.PARTICLE); p = rng.random((n, 3)) * 2 - 1
    if kind == "quant": p = np.round(p * 8) / 8
    elif kind == "line_x": p[:, 1:] = 0.5
    elif kind == "flat_y": p[:, 1] = 0.0
    elif kind == "two_scales": p[: n // 2] = 0.25 + rng.random((n // 2, 3)) * 1e-8
    elif kind == "dup": k = max(1, n // 3); p[k:2 * k] = p[:k][: len(p[k:2 * k])]
    elif kind == "grid":
        g = max(1, int(round(n ** (1 / 3)))); ids = rng.permutation(n)
        p = np.stack([ids % g, (ids // g) % g, ids // (g * g)], axis=1).astype(np.float64)
    parts["p"] = p; parts["v"] = rng.normal(size=(n, 3)) * 0.1
    parts["m"] = rng.random(n) / n
    if kind == "zero_mass": parts["m"][rng.random(n) < 0.3] = 0.0
    parts["r"] = 1e-3
    return parts
bad = 0
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    kind = KINDS[int(rng.integers(len(KINDS)))]
    world = int(rng.choice([2, 2, 3, 4, 5]))
    n = int(rng.choice([5, 40, 70, 130, 300, int(rng.integers(300, 6000))]))
    steps = int(rng.choice([1, 2, 3]))
    ics = make(kind, n, rng)
    uid = kd.KDTreeSim.comm_unique_id()
    out, errs = [None] * world, []
    ths = [threading.Thread(target=mr.rank_main, args=(r, world, uid, ics, steps, out, errs), daemon=True) for r in range(world)]
    [t.start() for t in ths]; [t.join(300) for t in ths]
    if errs or any(o is None for o in out):
        print(f"FAIL seed={seed} kind={kind} world={world} n={len(ics)}: {errs or 'hang'}", flush=True); os._exit(1)
    with kd.KDTreeSim() as one:
        one.upload(ics); one.build_tree(); one.calc_accel(); acc1 = one.accel(); one.simple_sim(1e-3, steps); res1 = one.download()
    def same(a, b):  # bit-identical incl. NaN payload positions
        return a.tobytes() == b.tobytes()
    ok = all(same(out[r][0], acc1) and same(out[r][1], res1) for r in range(world))
    if not ok:
        bad += 1; print(f"MISMATCH seed={seed} kind={kind} world={world} n={len(ics)} steps={steps}", flush=True)
    if (seed - first + 1) % 10 == 0: print(f"... {seed - first + 1} seeds, {bad} bad", flush=True)
print(f"{count} seeds from {first}: {bad} bad")
