"""Development aid: randomised parity campaign of the kernel SOURCES under the SIMT model (see run.py) against the
oracle — many small, awkward inputs (ties, duplicates, flat dimensions, collinear sets, mixed scales, signed zeros,
zero masses) across max_parts / layout / theta, checking what tests/test_gpu_parity.py checks:

  tree bit-exact (canonical order), every walk decision identical, accelerations <= 1e-11 relative (non-finite exactly where
  the reference is non-finite: coincident particles), a 3-step trajectory <= 1e-11 of the field's scale.

    python tests/devtools/simt/fuzz.py [first_seed] [count] [max_n]

Prints one line per failing seed and a summary; exit code 1 if any seed failed.  Not collected by pytest, never
touches the product (the ctypes loader of THIS process is re-pointed at the emulated library)."""
import os
import sys
import traceback

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

os.environ["KDNB_NO_GRAPH"] = "1"
import build as simt_build  # noqa: E402

lib = simt_build.build()
from multilanguagekdtree_b200 import _lib  # noqa: E402

_lib.LIB_PATH = lib
import multilanguagekdtree_b200 as kd  # noqa: E402
from oracle import okd  # noqa: E402
import test_gpu_parity as T  # noqa: E402  (helpers only: assert_tree_bit_exact, to_oracle_nodes)

okd.build()
ORC = {8: okd.Oracle(), 32: okd.Oracle(leaf_cap=32)}

KINDS = ["ring", "cube", "quant", "same_point", "line_x", "line_diag", "flat_x", "flat_y", "flat_z_lifted", "two_scales",
         "pairs", "dup_particles", "neg_zero", "tiny", "zero_mass", "ring_no_centre", "grid"]


def make(kind: str, n: int, rng: np.random.Generator) -> np.ndarray:
    parts = np.zeros(n, okd.PARTICLE)
    p = rng.random((n, 3)) * 2.0 - 1.0
    if kind == "ring" or kind == "ring_no_centre":
        parts = ORC[8].circular_orbits(max(1, n - 1), seed=int(rng.integers(1, 1 << 30)))
        if kind == "ring_no_centre" and len(parts) > 1:
            parts = parts[1:].copy()
        return parts
    if kind == "quant":
        q = int(rng.choice([2, 4, 16, 64]))
        p = np.round(p * q) / q
    elif kind == "same_point":
        p[:] = rng.random(3)
    elif kind == "line_x":
        p[:, 1:] = 0.5
    elif kind == "line_diag":
        p[:, 1] = p[:, 0]
        p[:, 2] = p[:, 0]
    elif kind == "flat_x":
        p[:, 0] = -0.25
    elif kind == "flat_y":
        p[:, 1] = 0.0
    elif kind == "flat_z_lifted":
        p[:, 2] = 0.125
    elif kind == "two_scales":
        p[: n // 2] = 0.25 + rng.random((n // 2, 3)) * 1e-8
    elif kind == "pairs":
        base = np.sort(rng.random((n + 1) // 2) * 2.0 - 1.0)
        x = np.empty(2 * len(base))
        x[0::2] = np.nextafter(base, 2.0)
        x[1::2] = base
        p[:, 0] = x[:n]
    elif kind == "dup_particles":
        k = max(1, n // 3)
        p[k:2 * k] = p[:k][: len(p[k:2 * k])]
    elif kind == "neg_zero":
        p[:, 2] = np.where(rng.random(n) < 0.5, 0.0, -0.0)
        p[rng.random(n) < 0.2, 0] = -0.0
    elif kind == "tiny":
        p *= 1e-300
    elif kind == "grid":
        g = max(1, int(round(n ** (1 / 3))))
        ids = np.arange(n)
        p = np.stack([ids % g, (ids // g) % g, ids // (g * g)], axis=1).astype(np.float64)
    parts["p"] = p
    parts["v"] = rng.normal(size=(n, 3)) * 0.1
    parts["m"] = rng.random(n) / n if rng.random() < 0.5 else 1.0 / n
    if kind == "zero_mass":
        parts["m"][rng.random(n) < 0.3] = 0.0
    parts["r"] = 1e-3
    return parts


def same_or_both_nonfinite(a, b, rtol):
    """Per-particle relative error where both are finite; the non-finite pattern must agree.  (Where a squared pair
    distance underflows to 0 the reference divides by zero: +-inf or NaN.  The default rsqrt path is non-finite in
    exactly the same places but may hold NaN where the reference holds inf; KDNB_FLAG_EXACT_MATH is bit-identical.)"""
    a, b = np.asarray(a), np.asarray(b)
    fa, fb = np.isfinite(a).all(axis=1), np.isfinite(b).all(axis=1)
    if not np.array_equal(fa, fb):
        return False, "non-finite pattern differs"
    aa, bb = a[fa], b[fa]
    if len(aa) == 0:
        return True, ""
    err = np.linalg.norm(aa - bb, axis=1) / np.maximum(np.linalg.norm(bb, axis=1), 1e-300)
    return bool(err.max() <= rtol), f"max rel err {err.max():.3e}"


def close_on_the_global_scale(a, b, rtol):
    """Trajectories: a close pair (|a| ~ 1e12 on collinear inputs) amplifies last-bit differences of one step into the
    next, so the comparison is against the largest component of the field, not per particle."""
    a, b = np.asarray(a), np.asarray(b)
    fa, fb = np.isfinite(a).all(axis=1), np.isfinite(b).all(axis=1)
    if not np.array_equal(fa, fb):
        return False, "non-finite pattern differs"
    if not fa.any():
        return True, ""
    scale = max(np.abs(b[fa]).max(), 1e-300)
    err = np.abs(a[fa] - b[fa]).max() / scale
    return bool(err <= rtol), f"max err / max |ref| = {err:.3e}"


def one(seed: int, max_n: int) -> str:
    rng = np.random.default_rng(seed)
    kind = KINDS[int(rng.integers(len(KINDS)))]
    n = int(rng.choice([1, 2, 3, 7, 8, 9, 16, 17, 31, 32, 33, 63, 64, 65, int(rng.integers(1, 300)),
                        int(rng.integers(300, max(301, max_n)))]))
    mp = int(rng.choice([4, 5, 7, 8, 8, 8, 16, 32]))
    layout = int(rng.choice([kd.LAYOUT_PADDED, kd.LAYOUT_PADDED, kd.LAYOUT_DENSE]))
    theta = float(rng.choice([0.2, 0.3, 0.3, 0.5, 0.7, 0.9]))
    orc = ORC[8 if mp <= 8 else 32]
    parts = make(kind, n, rng)
    tag = f"seed={seed} kind={kind} n={len(parts)} mp={mp} layout={layout} theta={theta}"
    olayout = okd.LAYOUT_PADDED if layout == kd.LAYOUT_PADDED else okd.LAYOUT_DENSE
    with kd.KDTreeSim(max_parts=mp, layout=layout, theta=theta, flags=kd.FLAG_WALK_COUNTS) as sim:
        sim.upload(parts)
        sim.build_tree()
        gnodes, gidx = sim.tree()
        onodes, oidx, _ = orc.build_tree_canonical(parts, max_parts=mp, layout=olayout)
        T.assert_tree_bit_exact(orc, gnodes, gidx, onodes, oidx, mp)
        sim.calc_accel()
        acc, cnt = sim.accel(), sim.walk_counts()
    on = T.to_oracle_nodes(orc, gnodes, gidx, mp)
    oacc, ocnt = orc.calc_accel_all(parts, on, theta=theta, counts=True)
    for k, f in enumerate(("node_visits", "accepts", "leaf_visits", "pp")):
        assert np.array_equal(cnt[:, k], ocnt[f]), f"{f} differ"
    ok, why = same_or_both_nonfinite(acc, oacc, 1e-11)
    assert ok, "acc: " + why
    # production walk kernel (no counters): bit-identical to the counting variant
    with kd.KDTreeSim(max_parts=mp, layout=layout, theta=theta) as sim:
        sim.upload(parts)
        sim.build_tree()
        sim.calc_accel()
        acc2 = sim.accel()
        assert np.array_equal(T.bits(acc2), T.bits(acc)), "production walk != counting walk"
        # short trajectory.  Where the first kick throws the whole set astronomically far (coincident particles whose
        # centre of mass is one ulp off: |a| ~ 1e31), the next step depends on whether the new positions coincide to
        # the last bit — the summation order decides between finite and NaN, in the reference as well — so only one
        # step is comparable.
        fin = np.isfinite(acc).all(axis=1)
        blown = bool(fin.any()) and float(np.abs(acc[fin]).max()) * 1e-6 > 1e3 * max(float(np.abs(parts["p"]).max()), 1e-300)
        steps = 1 if blown else 3
        sim.upload(parts)
        sim.simple_sim(1e-3, steps)
        out = sim.download()
    ref = parts.copy()
    orc.simple_sim(ref, 1e-3, steps, max_parts=mp, theta=theta, layout=olayout, order=okd.ORDER_CANONICAL)
    for f in ("p", "v"):
        ok, why = close_on_the_global_scale(out[f], ref[f], 1e-11)
        assert ok, f"trajectory {f}: {why}"
    return tag


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    max_n = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
    bad = 0
    for seed in range(first, first + count):
        try:
            one(seed, max_n)
        except Exception as e:  # noqa: BLE001
            bad += 1
            rng = np.random.default_rng(seed)
            kind = KINDS[int(rng.integers(len(KINDS)))]
            print(f"FAIL seed={seed} kind={kind}: {type(e).__name__}: {str(e)[:300]}")
            if os.environ.get("FUZZ_TRACE"):
                traceback.print_exc()
        if (seed - first + 1) % 25 == 0:
            print(f"... {seed - first + 1} seeds, {bad} failed", flush=True)
    print(f"{count} seeds from {first}: {bad} failed")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
