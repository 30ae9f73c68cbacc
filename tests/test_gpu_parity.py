"""GPU parity tests: the CUDA path, called through the C ABI (libkdnb.so), against the CPU oracle on the same inputs.

Contract:
  bit-exact : node index layout, kind, num_parts, leaf membership, indices, split_dim, split_val, size, left, right,
              and m / cm against the oracle's canonical summation order; kick/drift given identical accelerations;
              per-particle counts of node tests / accepts / leaf visits / pair interactions (i.e. every acceptance decision)
  tolerance : accelerations <= 1e-12 relative (vector norm) against the oracle's pairwise walk; positions after K steps
              <= 1e-12 relative; m / cm against the FAITHFUL (reference summation order) oracle <= 1e-11.
"""
import numpy as np
import pytest

import multilanguagekdtree_b200 as kd
from oracle.okd import LAYOUT_DENSE as O_DENSE
from oracle.okd import LAYOUT_PADDED as O_PADDED
from oracle.okd import ORDER_CANONICAL, ORDER_FAITHFUL, PARTICLE

pytestmark = pytest.mark.gpu

ACC_RTOL = 1e-12
POS_RTOL = 1e-12


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def to_oracle_nodes(orc, nodes, indices, max_parts):
    """kdnb_node records -> the oracle's okd_node records (for its walk / invariant checker)."""
    out = orc.allocate_node_vec(len(nodes))
    internal = nodes["kind"] == kd.INTERNAL
    out["is_internal"] = internal
    lp = kd.leaf_parts(nodes, indices, orc.leaf_cap)
    leaf = ~internal
    out["num_parts"][leaf] = nodes["num_parts"][leaf]
    out["leaf_parts"][leaf] = lp[leaf]
    for f in ("split_dim", "split_val", "m", "cm", "size", "left", "right"):
        out[f][internal] = nodes[f][internal]
    return out


def assert_tree_bit_exact(orc, gnodes, gidx, onodes, oidx, max_parts, sums_exact=True):
    assert len(gnodes) >= len(onodes) or len(onodes) >= len(gnodes)
    n = min(len(gnodes), len(onodes))
    g, o = gnodes[:n], onodes[:n]
    internal = o["is_internal"].astype(bool)
    assert np.array_equal(g["kind"] == kd.INTERNAL, internal)
    for f in ("split_dim", "left", "right"):
        assert np.array_equal(g[f][internal].astype(np.uint64), o[f][internal]), f
    for f in ("split_val", "size"):
        assert np.array_equal(g[f][internal], o[f][internal]), f          # value-exact (-0.0 == +0.0)
    leaf = ~internal
    assert np.array_equal(g["num_parts"][leaf], o["num_parts"][leaf])
    glp = kd.leaf_parts(g, gidx, orc.leaf_cap)
    used = leaf & (o["num_parts"] > 0)
    assert np.array_equal(glp[used], o["leaf_parts"][used])                # canonical order: ascending ids, 0 padding
    unused = leaf & (o["num_parts"] == 0)
    assert np.all(glp[unused] == np.uint64(kd.NO_INDEX)) and np.all(o["leaf_parts"][unused] == np.uint64(kd.NO_INDEX))
    assert np.array_equal(gidx, oidx)
    if sums_exact:
        assert np.array_equal(bits(g["m"][internal]), bits(o["m"][internal]))
        assert np.array_equal(bits(g["cm"][internal]), bits(o["cm"][internal]))


def cube(n, seed, quant=None, equal_mass=True):
    rng = np.random.default_rng(seed)
    parts = np.zeros(n, PARTICLE)
    p = rng.random((n, 3)) * 2.0 - 1.0
    if quant:
        p = np.round(p * quant) / quant       # many exactly equal coordinates: exercises the tie-break
    parts["p"] = p
    parts["v"] = rng.normal(size=(n, 3)) * 0.1
    parts["m"] = 1.0 / n if equal_mass else rng.random(n) / n
    parts["r"] = 1e-3
    return parts


@pytest.mark.parametrize("n", [1, 2, 8, 9, 11, 16, 17, 100, 1000, 2047, 2048, 2049, 5000, 4096 * 3 + 5, 100000])
def test_build_padded_ring_bit_exact(orc, n):
    parts = orc.circular_orbits(n, seed=1000 + n)
    with kd.KDTreeSim() as sim:
        sim.upload(parts)
        sim.build_tree()
        gnodes, gidx = sim.tree()
    onodes, oidx, _ = orc.build_tree_canonical(parts, layout=O_PADDED, threads=8)
    assert len(gnodes) == len(onodes) == orc.nodes_needed_for_particles(n + 1, 8)
    assert_tree_bit_exact(orc, gnodes, gidx, onodes, oidx, 8)
    assert orc.check_tree_struct(to_oracle_nodes(orc, gnodes, gidx, 8), parts) == 0   # array_kd_tree.rs:834-877


@pytest.mark.parametrize("n", [3, 12, 500, 5001, 70001])
def test_build_dense_bit_exact(orc, n):
    parts = orc.circular_orbits(n - 1, seed=n)
    with kd.KDTreeSim(layout=kd.LAYOUT_DENSE) as sim:
        sim.upload(parts)
        sim.build_tree()
        gnodes, gidx = sim.tree()
    onodes, oidx, last = orc.build_tree_canonical(parts, layout=O_DENSE)
    assert len(gnodes) == last + 1
    assert_tree_bit_exact(orc, gnodes, gidx, onodes[: last + 1], oidx, 8)


@pytest.mark.parametrize("mp,layout", [(7, kd.LAYOUT_DENSE), (4, kd.LAYOUT_PADDED), (5, kd.LAYOUT_DENSE)])
def test_build_other_max_parts(orc, mp, layout):
    parts = orc.circular_orbits(20000, seed=mp)
    with kd.KDTreeSim(max_parts=mp, layout=layout) as sim:
        sim.upload(parts)
        sim.build_tree()
        gnodes, gidx = sim.tree()
    onodes, oidx, last = orc.build_tree_canonical(parts, max_parts=mp, layout=O_PADDED if layout == kd.LAYOUT_PADDED else O_DENSE)
    assert_tree_bit_exact(orc, gnodes, gidx, onodes[: len(gnodes)], oidx, mp)


@pytest.mark.parametrize("mp", [16, 32])
def test_build_large_leaves(orc32, mp):
    parts = orc32.circular_orbits(30000, seed=mp)
    with kd.KDTreeSim(max_parts=mp) as sim:
        sim.upload(parts)
        sim.build_tree()
        gnodes, gidx = sim.tree()
    onodes, oidx, _ = orc32.build_tree_canonical(parts, max_parts=mp)
    assert_tree_bit_exact(orc32, gnodes, gidx, onodes, oidx, mp)


@pytest.mark.parametrize("n,quant", [(5000, None), (40000, None), (30000, 64), (3000, 4)])
def test_build_3d_and_ties_bit_exact(orc, n, quant):
    parts = cube(n, seed=n, quant=quant, equal_mass=False)
    with kd.KDTreeSim() as sim:
        sim.upload(parts)
        sim.build_tree()
        gnodes, gidx = sim.tree()
    onodes, oidx, _ = orc.build_tree_canonical(parts, threads=4)
    assert_tree_bit_exact(orc, gnodes, gidx, onodes, oidx, 8)
    assert set(np.unique(gnodes["split_dim"][gnodes["kind"] == kd.INTERNAL])) == {0, 1, 2}


def _clustered(n, seed, kind):
    """Inputs that defeat the 32-bit sort keys (sort.cu): far more distinct coordinates than 2^-32 of the extent
    can separate, so whole runs share a key32 and the build has to fall back on (or fix up with) the full 64-bit order."""
    rng = np.random.default_rng(seed)
    parts = cube(n, seed=seed, equal_mass=False)
    if kind == "tight_cluster_far_outlier":        # all but one particle within 1e-9 of each other, one at 1e3
        parts["p"] = 1.0 + rng.random((n, 3)) * 1e-9
        parts["p"][n // 2] = (1e3, -1e3, 5e2)
    elif kind == "two_scales":                     # half the particles in a 1e-8 ball, half spread over [-1, 1)
        parts["p"][: n // 2] = 0.25 + rng.random((n // 2, 3)) * 1e-8
    elif kind == "huge_extent":                    # hi - lo overflows to inf
        parts["p"][0] = (1.5e308, 1.5e308, 1.5e308)
        parts["p"][1] = (-1.5e308, -1.5e308, -1.5e308)
    elif kind == "pairs":                          # neighbours one ulp apart, in descending id order
        base = np.sort(rng.random(n // 2) * 2.0 - 1.0)
        x = np.empty(n)
        x[0::2] = np.nextafter(base, 2.0)
        x[1::2] = base
        parts["p"][:, 0] = x
    return parts


@pytest.mark.parametrize("kind", ["tight_cluster_far_outlier", "two_scales", "huge_extent", "pairs"])
def test_build_inputs_beyond_32_bit_keys_bit_exact(orc, kind):
    parts = _clustered(30000, seed=11, kind=kind)
    with kd.KDTreeSim() as sim:
        sim.upload(parts)
        sim.build_tree()
        gnodes, gidx = sim.tree()
    onodes, oidx, _ = orc.build_tree_canonical(parts, threads=4)
    assert_tree_bit_exact(orc, gnodes, gidx, onodes, oidx, 8)


def test_build_forced_64_bit_sort_matches():
    """KDNB_SORT=64 skips the 32-bit passes, KDNB_SORT=stubs launches gated 64-bit passes: same tree every way."""
    import subprocess, sys
    code = ("import numpy as np, multilanguagekdtree_b200 as kd\n"
            "p = kd.circular_orbits(50000, seed=5)\n"
            "s = kd.KDTreeSim(); s.upload(p); s.build_tree(); n, i = s.tree()\n"
            "np.save(__import__('sys').argv[1], i)\n")
    import tempfile, os
    with tempfile.TemporaryDirectory() as d:
        out = {}
        for tag, env in (("k32", {}), ("k64", {"KDNB_SORT": "64"}), ("stubs", {"KDNB_SORT": "stubs"})):
            f = os.path.join(d, tag + ".npy")
            subprocess.run([sys.executable, "-c", code, f], check=True, env={**os.environ, **env},
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            out[tag] = np.load(f)
        assert np.array_equal(out["k32"], out["k64"]) and np.array_equal(out["k32"], out["stubs"])


def test_walk_launch_order_does_not_change_results():
    """The production walk hands out its 32-particle groups heaviest first (previous step's work, walk_order_kernel).
    Which CTA computes which group must not matter: accelerations of the 2nd and 3rd walk (the ordered ones) and the
    state after 4 steps are bit-identical with KDNB_WALK_LPT=0 (index order)."""
    import os
    import subprocess
    import sys
    import tempfile
    code = ("import sys, numpy as np, multilanguagekdtree_b200 as kd\n"
            "p = kd.circular_orbits(30000, seed=8)\n"
            "s = kd.KDTreeSim(); s.upload(p); acc = []\n"
            "for k in range(3):\n"
            "    s.build_tree(); s.calc_accel(); acc.append(s.accel()); s.kick_drift(1e-3)\n"
            "s.simple_sim(1e-3, 4)\n"
            "np.savez(sys.argv[1], acc=np.stack(acc), out=s.download())\n")
    with tempfile.TemporaryDirectory() as d:
        res = {}
        for tag, env in (("lpt", {}), ("index", {"KDNB_WALK_LPT": "0"})):
            f = os.path.join(d, tag + ".npz")
            subprocess.run([sys.executable, "-c", code, f], check=True, env={**os.environ, **env},
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            res[tag] = np.load(f)
        assert np.array_equal(bits(res["lpt"]["acc"]), bits(res["index"]["acc"]))
        assert res["lpt"]["out"].tobytes() == res["index"]["out"].tobytes()


def test_build_vs_faithful_reference_order(orc):
    """Against the reference's own summation order (random pivots): everything order-independent is bit-exact,
    m / cm agree to the reference's run-to-run noise."""
    parts = orc.circular_orbits(50000, seed=77)
    with kd.KDTreeSim() as sim:
        sim.upload(parts)
        sim.build_tree()
        g, gidx = sim.tree()
    o, oidx = orc.build_tree_par4(parts, seed=5, threads=4)
    internal = o["is_internal"].astype(bool)
    assert np.array_equal(g["kind"] == kd.INTERNAL, internal)
    for f in ("split_dim", "left", "right"):
        assert np.array_equal(g[f][internal].astype(np.uint64), o[f][internal])
    for f in ("split_val", "size"):
        assert np.array_equal(g[f][internal], o[f][internal])
    glp = kd.leaf_parts(g, gidx)
    leaf = ~internal & (o["num_parts"] > 0)
    assert np.array_equal(g["num_parts"][leaf], o["num_parts"][leaf])
    assert np.array_equal(np.sort(glp[leaf], axis=1), np.sort(o["leaf_parts"][leaf], axis=1))   # leaf membership as sets
    assert np.allclose(g["m"][internal], o["m"][internal], rtol=1e-11, atol=0)
    assert np.allclose(g["cm"][internal], o["cm"][internal], rtol=0, atol=1e-11)


def _walk_case(orc, parts, flags=0, theta=0.3, mp=8):
    with kd.KDTreeSim(flags=flags | kd.FLAG_WALK_COUNTS, theta=theta, max_parts=mp) as sim:
        sim.upload(parts)
        sim.build_tree()
        sim.calc_accel()
        acc = sim.accel()
        cnt = sim.walk_counts()
        gnodes, gidx = sim.tree()
    onodes = to_oracle_nodes(orc, gnodes, gidx, mp)
    oacc, ocnt = orc.calc_accel_all(parts, onodes, theta=theta, counts=True)
    return acc, cnt, oacc, ocnt


def rel_err(a, b):
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)


@pytest.mark.parametrize("n", [1, 11, 1000, 5000, 100000])
def test_walk_ring_acc_and_decisions(orc, n):
    parts = orc.circular_orbits(n, seed=n + 5)
    acc, cnt, oacc, ocnt = _walk_case(orc, parts)
    for k, f in enumerate(("node_visits", "accepts", "leaf_visits", "pp")):
        assert np.array_equal(cnt[:, k], ocnt[f]), f       # every open/accept decision identical to the reference's
    assert rel_err(acc, oacc).max() <= ACC_RTOL


def test_one_context_walks_different_particle_counts(orc):
    """Per-size host state of the walk (the closed-form node indices of the top of the tree that seed the traversal,
    kdnb_api.cu: plan) must follow the particle count when one context is loaded with another set: growing, shrinking to
    a size without that seed (a tree shallower than five internal levels), and back."""
    with kd.KDTreeSim(flags=kd.FLAG_WALK_COUNTS) as sim:
        for n in (3000, 7000, 60, 3000, 1200):
            parts = orc.circular_orbits(n, seed=n + 11)
            sim.upload(parts)
            sim.build_tree()
            sim.calc_accel()
            acc, cnt = sim.accel(), sim.walk_counts()
            gnodes, gidx = sim.tree()
            oacc, ocnt = orc.calc_accel_all(parts, to_oracle_nodes(orc, gnodes, gidx, 8), counts=True)
            for k, f in enumerate(("node_visits", "accepts", "leaf_visits", "pp")):
                assert np.array_equal(cnt[:, k], ocnt[f]), (n, f)
            assert rel_err(acc, oacc).max() <= ACC_RTOL, n


def test_walk_ring_self_gravity_visible(orc):
    """SURVEY.md §7 hard part 4: the central mass (m=1 vs 1e-14) hides ring-ring errors at the 1e-12 level.
    Drop the central body so that every force IS ring self-gravity and opening mistakes would show."""
    parts = orc.circular_orbits(20000, seed=3)[1:].copy()
    acc, cnt, oacc, ocnt = _walk_case(orc, parts)
    for k, f in enumerate(("node_visits", "accepts", "leaf_visits", "pp")):
        assert np.array_equal(cnt[:, k], ocnt[f]), f
    assert rel_err(acc, oacc).max() <= 1e-11


@pytest.mark.parametrize("theta", [0.2, 0.3, 0.5, 0.7])
def test_walk_equal_mass_cube(orc, theta):
    parts = cube(20000, seed=9)
    acc, cnt, oacc, ocnt = _walk_case(orc, parts, theta=theta)
    for k, f in enumerate(("node_visits", "accepts", "leaf_visits", "pp")):
        assert np.array_equal(cnt[:, k], ocnt[f]), f
    assert rel_err(acc, oacc).max() <= 1e-11     # heavy cancellation in a uniform cube: |sum| << sum|terms|


def _acc(parts, flags=0, mp=8, theta=0.3):
    with kd.KDTreeSim(flags=flags, theta=theta, max_parts=mp) as sim:
        sim.upload(parts)
        sim.build_tree()
        sim.calc_accel()
        return sim.accel()


def _fast_path_cases(orc):
    ring = orc.circular_orbits(30000, seed=77)                       # planar, z == +0, masses > 0: z terms skipped
    neg0 = ring.copy(); neg0["p"][::3, 2] = -0.0                      # -0.0 is still planar
    lifted = ring.copy(); lifted["p"][:, 2] = 0.25                    # flat z list but cm_z != z exactly: general path
    massless = ring.copy(); massless["m"][5] = 0.0                    # a zero mass disables the planar shortcut
    return {"ring": ring, "neg0": neg0, "lifted": lifted, "massless": massless,
            "cube": cube(20000, seed=4), "cube_unequal": cube(20000, seed=5, equal_mass=False)}


@pytest.mark.parametrize("case", ["ring", "neg0", "lifted", "massless", "cube", "cube_unequal"])
def test_walk_production_kernel_equals_counted_kernel(orc, case):
    """The parity tests above run the counting variant of the walk kernel; the variant that simple_sim launches
    (no counters, z terms skipped for planar inputs) must give the same accelerations BIT FOR BIT: same traversal,
    same interaction lists, same operation order (a skipped z term is exactly +-0)."""
    parts = _fast_path_cases(orc)[case]
    fast = _acc(parts)
    counted = _acc(parts, flags=kd.FLAG_WALK_COUNTS)
    assert np.array_equal(fast, counted)        # value-exact (-0.0 == +0.0)
    assert np.all(np.isfinite(fast))
    if case in ("ring", "neg0"):
        assert np.all(fast[:, 2] == 0.0)


@pytest.mark.parametrize("shape", ["ring", "cube"])
def test_walk_production_kernel_on_a_grid_of_several_waves(orc, shape):
    """From two waves of 28 x (number of SMs) one-warp CTAs the production walk is launched in its 72-register build
    (walk.cu: launch_walk2) — sizes the other walk tests never reach.  Same source, other register allocation: it must
    still equal the counting variant bit for bit, planar (z terms skipped, 120-entry list) and general."""
    parts = orc.circular_orbits(320_000, seed=91) if shape == "ring" else cube(300_000, seed=92, equal_mass=False)
    fast = _acc(parts)
    counted = _acc(parts, flags=kd.FLAG_WALK_COUNTS)
    assert np.array_equal(fast, counted)
    assert np.all(np.isfinite(fast))


@pytest.mark.parametrize("mp", [4, 7])
@pytest.mark.parametrize("shape", ["ring", "cube"])
def test_walk_other_max_parts(orc, mp, shape):
    parts = orc.circular_orbits(20000, seed=mp) if shape == "ring" else cube(12000, seed=mp)
    acc, cnt, oacc, ocnt = _walk_case(orc, parts, mp=mp)
    for k, f in enumerate(("node_visits", "accepts", "leaf_visits", "pp")):
        assert np.array_equal(cnt[:, k], ocnt[f]), f
    assert rel_err(acc, oacc).max() <= (ACC_RTOL if shape == "ring" else 1e-11)
    assert np.array_equal(_acc(parts, mp=mp), acc)


@pytest.mark.parametrize("mp", [16, 32])
def test_walk_large_leaves(orc32, mp):
    parts = orc32.circular_orbits(20000, seed=mp)
    acc, cnt, oacc, ocnt = _walk_case(orc32, parts, mp=mp)
    for k, f in enumerate(("node_visits", "accepts", "leaf_visits", "pp")):
        assert np.array_equal(cnt[:, k], ocnt[f]), f
    assert rel_err(acc, oacc).max() <= ACC_RTOL
    assert np.array_equal(_acc(parts, mp=mp), acc)


def test_walk_exact_math_flag(orc):
    parts = cube(8000, seed=2)
    acc, _, oacc, _ = _walk_case(orc, parts, flags=kd.FLAG_EXACT_MATH)
    assert rel_err(acc, oacc).max() <= 1e-13     # same sqrt / divide as the reference; only the summation order differs


def test_kick_drift_bit_exact(orc):
    parts = orc.circular_orbits(30000, seed=8)
    rng = np.random.default_rng(0)
    acc = rng.normal(size=(len(parts), 3))
    with kd.KDTreeSim() as sim:
        sim.upload(parts)
        sim.build_tree()
        sim.set_accel(acc)
        assert np.array_equal(bits(sim.accel()), bits(acc))
        sim.kick_drift(1e-3)
        out = sim.download()
        assert not sim.accel().any()          # a[k] = 0 (array_kd_tree.rs:659-661)
    ref = parts.copy()
    orc.kick_drift(ref, acc.copy(), 1e-3)
    assert out.tobytes() == ref.tobytes()


def test_upload_download_roundtrip(orc):
    parts = cube(12345, seed=4)
    with kd.KDTreeSim() as sim:
        sim.upload(parts)
        assert sim.download().tobytes() == parts.tobytes()


@pytest.mark.parametrize("n,steps", [(1000, 100), (20000, 10)])
def test_simple_sim_trajectory(orc, n, steps):
    parts = orc.circular_orbits(n, seed=n)
    g = parts.copy()
    kd.simple_sim(g, 1e-3, steps)
    for order in (ORDER_CANONICAL, ORDER_FAITHFUL):
        o = parts.copy()
        orc.simple_sim(o, 1e-3, steps, order=order, seed=3)
        scale = np.abs(o["p"]).max()
        assert np.linalg.norm(g["p"] - o["p"], axis=1).max() / scale <= POS_RTOL
        assert np.linalg.norm(g["v"] - o["v"], axis=1).max() / np.abs(o["v"]).max() <= POS_RTOL
    assert np.array_equal(g["m"], parts["m"]) and np.array_equal(g["r"], parts["r"])


def test_simple_sim_sequential_crate_semantics(orc):
    """BASELINE config #1: Sequential/RustVersion (MAX_PARTS=7, dense layout), N=1000, 100 steps."""
    parts = orc.circular_orbits(1000, seed=1)
    g = parts.copy()
    with kd.KDTreeSim(max_parts=7, layout=kd.LAYOUT_DENSE) as sim:
        sim.simple_sim_bodies(g, 1e-3, 100)
    o = parts.copy()
    orc.simple_sim(o, 1e-3, 100, max_parts=7, layout=O_DENSE, order=ORDER_FAITHFUL, seed=9, threads=1)
    assert np.linalg.norm(g["p"] - o["p"], axis=1).max() / np.abs(o["p"]).max() <= POS_RTOL


def test_two_bodies_orbit(orc):
    """Parallel/GoVersion/kdtree_test.go:63-74 (prints only there): half orbit after 1000 steps of pi/1000."""
    g = kd.two_bodies()
    kd.simple_sim(g, np.pi / 1000, 1000)
    gold = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "pure_two_bodies.npz"))
    after = gold["after"].view(PARTICLE).reshape(-1)
    assert np.abs(g["p"] - after["p"]).max() <= 1e-12      # against the reference Python's own output
    assert abs(g["p"][1][0] + 1.0) < 2e-2


@pytest.mark.parametrize("name", ["pure_ring300_mp8", "pure3d_ring400_mp8"])
def test_golden_reference_python_trajectory(name):
    """The reference's own PureVersion output (tests/golden): accelerations and a few steps; the 3-D case (z chosen as
    split dimension in many nodes, see tests/golden/make_golden.py) also pins the dense tree's split planes bit for bit."""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    parts = gold["parts0"].view(PARTICLE).reshape(-1).copy()
    acc = kd.calc_accel_all(parts)
    assert rel_err(acc, gold["acc"]).max() <= ACC_RTOL
    # build_tree (dense layout) against the reference's own nodes: topology and split planes are order-independent
    nodes = kd.allocate_node_vec(len(parts))
    idx = np.arange(len(parts), dtype=np.uint64)
    last, nodes = kd.build_tree(idx, 0, len(parts), parts, 0, nodes)
    assert last == int(gold["last"])
    internal = gold["tree_is_internal"].astype(bool)
    t = nodes[: last + 1]
    assert np.array_equal(t["kind"] == kd.INTERNAL, internal)
    for f in ("split_dim", "left", "right"):
        assert np.array_equal(t[f][internal].astype(np.uint64), gold["tree_" + f][internal]), f
    for f in ("split_val", "size"):
        assert np.array_equal(t[f][internal].view(np.uint64), gold["tree_" + f][internal].view(np.uint64)), f
    kd.simple_sim(parts, float(gold["dt"]), int(gold["steps"]))
    after = gold["after"].view(PARTICLE).reshape(-1)
    assert np.abs(parts["p"] - after["p"]).max() / np.abs(after["p"]).max() <= POS_RTOL


def test_simd_particle_surface_equals_scalar_path():
    """Sequential crate, SIMD variant (simd_particle.rs:3-8, simd_kd_tree.rs:169-202): 96-byte f64x4 records through
    kdnb_*_simd give bit for bit what the scalar records give with the same constants (MAX_PARTS=7, dense layout);
    lane 3 comes back 0; a non-zero lane 3 is rejected."""
    from multilanguagekdtree_b200 import simd_kd_tree, simd_particle
    b = simd_particle.circular_orbits(3000, seed=5)
    b["p"][1:, 2] = 0.01 * np.sin(np.arange(1, len(b)))      # some z extent: all three lanes carry data
    s = simd_particle.to_scalar(b)
    simd_kd_tree.simple_sim(b, 1e-3, 4)
    kd.simple_sim(s, 1e-3, 4, max_parts=7)                   # scalar records, same constants (padded layout: same physics)
    assert np.array_equal(b["p"][:, :3].view(np.uint64), s["p"].view(np.uint64))
    assert np.array_equal(b["v"][:, :3].view(np.uint64), s["v"].view(np.uint64))
    assert not b["p"][:, 3].any() and not b["v"][:, 3].any()
    assert np.array_equal(b["m"], s["m"]) and np.array_equal(b["r"], s["r"])
    two = simd_particle.two_bodies()
    simd_kd_tree.simple_sim(two, np.pi / 1000, 1000)
    assert abs(two["p"][1][0] + 1.0) < 2e-2                  # half an orbit
    bad = simd_particle.circular_orbits(100, seed=1)
    bad["v"][7, 3] = 1e-3
    with pytest.raises(kd.KdnbError, match="lane 3"):
        simd_kd_tree.simple_sim(bad, 1e-3, 1)


def test_reference_api_mirror(orc):
    """two_leaves (array_kd_tree.rs:711-731) through the reference-shaped free functions."""
    parts = kd.circular_orbits(11)
    nodes = kd.allocate_node_vec(len(parts))
    idx = np.arange(len(parts), dtype=np.uint64)
    kd.build_tree_par4(idx, 0, parts, nodes, 1)
    assert nodes[0]["kind"] == kd.INTERNAL and nodes[1]["kind"] == kd.LEAF and nodes[2]["kind"] == kd.LEAF
    assert nodes[1]["num_parts"] + nodes[2]["num_parts"] == 12
    assert sorted(idx.tolist()) == list(range(12))
    # single_node (array_kd_tree.rs:698-709)
    two = kd.two_bodies()
    nodes = kd.allocate_node_vec(2)
    idx = np.arange(2, dtype=np.uint64)
    last, nodes = kd.build_tree(idx, 0, 2, two, 0, nodes)
    assert last == 0 and nodes[0]["kind"] == kd.LEAF and nodes[0]["num_parts"] == 2


def _degenerate(kind, n, seed):
    """Geometries that collapse one or two dimensions or put many particles on exactly equal coordinates."""
    rng = np.random.default_rng(seed)
    parts = cube(n, seed=seed, equal_mass=False)
    p = parts["p"].copy()
    if kind == "line_x":            # collinear along x: y and z lists are flat
        p[:, 1:] = 0.5
    elif kind == "line_diag":       # collinear along the diagonal: three identical orders, widest-axis ties
        p[:, 1] = p[:, 0]
        p[:, 2] = p[:, 0]
    elif kind == "flat_x":          # the x extent is 0: dimension 0 is never the widest
        p[:, 0] = -0.25
    elif kind == "flat_y":
        p[:, 1] = 0.0
    elif kind == "grid":            # integer lattice: every coordinate value shared by ~n^(2/3) particles
        g = max(1, int(round(n ** (1 / 3))))
        ids = rng.permutation(n)
        p = np.stack([ids % g, (ids // g) % g, ids // (g * g)], axis=1).astype(np.float64)
    parts["p"] = p
    return parts


@pytest.mark.parametrize("kind", ["line_x", "line_diag", "flat_x", "flat_y", "grid"])
@pytest.mark.parametrize("mp,layout", [(8, kd.LAYOUT_PADDED), (7, kd.LAYOUT_DENSE)])
def test_degenerate_geometries_tree_and_walk(orc, kind, mp, layout):
    parts = _degenerate(kind, 3000, seed=21)
    olayout = O_PADDED if layout == kd.LAYOUT_PADDED else O_DENSE
    with kd.KDTreeSim(max_parts=mp, layout=layout, flags=kd.FLAG_WALK_COUNTS) as sim:
        sim.upload(parts)
        sim.build_tree()
        gnodes, gidx = sim.tree()
        sim.calc_accel()
        acc, cnt = sim.accel(), sim.walk_counts()
    onodes, oidx, _ = orc.build_tree_canonical(parts, max_parts=mp, layout=olayout)
    assert_tree_bit_exact(orc, gnodes, gidx, onodes, oidx, mp)
    oacc, ocnt = orc.calc_accel_all(parts, to_oracle_nodes(orc, gnodes, gidx, mp), counts=True)
    for k, f in enumerate(("node_visits", "accepts", "leaf_visits", "pp")):
        assert np.array_equal(cnt[:, k], ocnt[f]), f
    # On a line or a lattice the pulls from both sides nearly cancel (|sum| down to 1e-4 of the sum of |terms|), so the
    # summation-order difference (running sum here, pairwise in the reference) shows at 1e-11; decisions are exact.
    assert np.isfinite(acc).all() and rel_err(acc, oacc).max() <= 1e-9


def test_empty_bodies_are_a_no_op_like_the_reference(orc):
    """`simple_sim(&mut vec![], dt, steps)` does not panic in the reference: acc / indices are empty, the tree is the
    one node of allocate_node_vec(0) (array_kd_tree.rs:45-60) which the build writes as Leaf{0, [0; MAX_PARTS]}
    (:524-529), and every stage loops over nothing."""
    empty = np.zeros(0, kd.PARTICLE)
    onodes, oidx, olast = orc.build_tree_canonical(np.zeros(0, PARTICLE))
    assert len(onodes) == 1 and olast == 0 and not onodes[0]["is_internal"] and onodes[0]["num_parts"] == 0
    with kd.KDTreeSim(flags=kd.FLAG_WALK_COUNTS) as sim:
        sim.upload(empty)
        assert sim.count == 0 and sim.node_count == 1
        with pytest.raises(kd.KdnbError):
            sim.calc_accel()                     # call order is still checked
        sim.build_tree()
        nodes, idx = sim.tree()
        assert len(nodes) == 1 and len(idx) == 0
        assert nodes[0]["kind"] == kd.LEAF and nodes[0]["num_parts"] == 0
        assert np.array_equal(kd.leaf_parts(nodes, idx)[0], onodes[0]["leaf_parts"])   # 0 padding, not usize::MAX
        sim.calc_accel()
        assert sim.accel().shape == (0, 3) and sim.walk_counts().shape == (0, 4)
        sim.kick_drift(1e-3)
        sim.simple_sim(1e-3, 5)
        assert len(sim.download()) == 0
        # the context is reusable afterwards, and an empty upload after a real one resets it
        parts = kd.circular_orbits(100)
        sim.upload(parts)
        sim.build_tree()
        assert sim.count == 101 and sim.node_count == kd.nodes_needed_for_particles(101)
        sim.upload(empty)
        assert sim.count == 0 and sim.node_count == 1
        with pytest.raises(kd.KdnbError):
            sim.calc_accel()                     # the tree of the previous upload is gone
    kd.simple_sim(empty, 1e-3, 3)                # the one-call form


def test_error_codes_instead_of_panics():
    with kd.KDTreeSim() as sim:
        with pytest.raises(kd.KdnbError):
            sim.build_tree()                     # nothing uploaded
        sim.upload(kd.circular_orbits(100))
        with pytest.raises(kd.KdnbError):
            sim.calc_accel()                     # no tree yet
        sim.build_tree()
        sim.calc_accel()
        sim.kick_drift(1e-3)
        with pytest.raises(kd.KdnbError):
            sim.calc_accel()                     # tree is stale after the drift
    with pytest.raises(kd.KdnbError):
        kd.KDTreeSim(max_parts=3)


def test_big_solar_with_steps_invariant(orc):
    """array_kd_tree.rs:816-832 — 10 steps, rebuild, the partition invariant still holds."""
    parts = orc.circular_orbits(5000, seed=6)
    with kd.KDTreeSim() as sim:
        sim.upload(parts)
        sim.simple_sim(1e-3, 10)
        sim.build_tree()
        nodes, idx = sim.tree()
        moved = sim.download()
    assert orc.check_tree_struct(to_oracle_nodes(orc, nodes, idx, 8), moved) == 0


def test_full_size_properties_1m(orc):
    """BASELINE config #3 size (N=1,000,000): size-independent properties + a sampled oracle comparison."""
    n = 1_000_000
    parts = orc.circular_orbits(n, seed=12345)
    with kd.KDTreeSim(flags=kd.FLAG_WALK_COUNTS) as sim:
        sim.upload(parts)
        sim.build_tree()
        nodes, idx = sim.tree()
        sim.calc_accel()
        acc = sim.accel()
        cnt = sim.walk_counts()
    assert len(nodes) == 524287
    assert np.array_equal(np.sort(idx), np.arange(n + 1, dtype=np.uint64))              # indices is a permutation
    used = nodes[(nodes["kind"] == kd.INTERNAL) | (nodes["num_parts"] > 0)]
    assert len(used) == 262143 and nodes["num_parts"].sum() == n + 1
    onodes = to_oracle_nodes(orc, nodes, idx, 8)
    assert orc.check_tree_struct(onodes, parts) == 0
    # tree bit-exact against the canonical oracle at full size
    cn, cidx, _ = orc.build_tree_canonical(parts, threads=8)
    assert_tree_bit_exact(orc, nodes, idx, cn, cidx, 8)
    # walk: all particles against the oracle (OpenMP), decisions exact
    oacc, ocnt = orc.calc_accel_all(parts, onodes, counts=True)
    assert rel_err(acc, oacc).max() <= ACC_RTOL
    for k, f in enumerate(("node_visits", "accepts", "leaf_visits", "pp")):
        assert np.array_equal(cnt[:, k], ocnt[f]), f


def test_print_tree_format_matches_oracle_dump(orc, tmp_path):
    """array_kd_tree.rs:666-692 / TreeVisualizer reader: same records as the oracle's dump of the same (canonical) tree."""
    parts = orc.circular_orbits(500, seed=13)
    with kd.KDTreeSim() as sim:
        sim.upload(parts)
        sim.build_tree()
        nodes, idx = sim.tree()
    gpath = kd.print_tree(0, nodes, idx, parts, str(tmp_path))
    onodes, _, _ = orc.build_tree_canonical(parts)
    opath = str(tmp_path / "oracle_tree0.txt")
    orc.print_tree(opath, onodes, parts)
    g, o = open(gpath).read().split("\n"), open(opath).read().split("\n")
    assert len(g) == len(o) and g[0] == o[0] == str(len(nodes))
    for lg, lo in zip(g[1:], o[1:]):
        tg, to = lg.split(), lo.split()
        assert len(tg) == len(to)
        for a, b in zip(tg, to):
            if a in ("L", "I"):
                assert a == b
            else:
                assert float(a) == float(b)          # %.17g there, Rust-style shortest form here: equal as numbers
                assert "e" not in a.lower()            # Rust's `{}` never prints an exponent


def test_graph_replay_equals_plain_launches(orc):
    """kdnb_simple_sim replays the step as a CUDA graph for calls of >= 3 steps; shorter calls use plain launches.
    Both must give bit-identical trajectories."""
    parts = orc.circular_orbits(30000, seed=5)
    with kd.KDTreeSim() as a, kd.KDTreeSim() as b:
        a.upload(parts)
        b.upload(parts)
        a.simple_sim(1e-3, 8)            # 1 plain step + 7 graph replays
        for _ in range(4):
            b.simple_sim(1e-3, 2)        # plain launches only
        assert a.download().tobytes() == b.download().tobytes()
        a.simple_sim(1e-3, 5)            # graph reused
        b.simple_sim(1e-3, 2); b.simple_sim(1e-3, 2); b.simple_sim(1e-3, 1)
        assert a.download().tobytes() == b.download().tobytes()


def test_simple_sim_zero_steps_and_staged_semantics(orc):
    """`for _ in 0..steps` (array_kd_tree.rs:632): steps <= 0 never advances the state — also on the call that would
    capture the step graph (third consecutive call).  And the reference's `a = 0` after a kick (:659-661): a second kick
    without a walk in between applies zero accelerations, and downloading consumed accelerations gives zeros."""
    parts = orc.circular_orbits(5000, seed=8)
    with kd.KDTreeSim() as sim:
        sim.upload(parts)
        for steps in (0, -1, 0, 0, -5):
            sim.simple_sim(1e-3, steps)
        assert sim.download().tobytes() == parts.tobytes()
        sim.build_tree()
        assert not sim.accel().any()                 # before the first calc_accel
        sim.calc_accel()
        acc = sim.accel()
        assert acc.any()
        sim.kick_drift(1e-3)
        after1 = sim.download()
        assert not sim.accel().any()                 # consumed: a = 0
        sim.kick_drift(1e-3)                         # no walk in between: v unchanged, p += dt * v
        after2 = sim.download()
    v1 = parts["v"] + 1e-3 * acc
    assert np.array_equal(after1["v"], v1) and np.array_equal(after2["v"], v1)
    assert np.array_equal(after2["p"], after1["p"] + 1e-3 * v1)


@pytest.mark.parametrize("kind", ["two_scales", "pairs"])
def test_graph_replay_with_64_bit_sort_fallback(orc, kind):
    """Inside a replayed step the 64-bit sort passes sit in a conditional graph node; inputs that need them (and
    inputs that only need the fix-up) must give the plain-launch trajectory bit for bit, and the oracle's within 1e-12."""
    parts = _clustered(20000, seed=3, kind=kind)
    parts["v"] *= 1e-6                   # the tight cluster stays tight: every step needs the same sort path
    parts["m"] *= 1e-20                  # ... and close pairs do not blow the trajectory up
    with kd.KDTreeSim() as a, kd.KDTreeSim() as b:
        a.upload(parts)
        b.upload(parts)
        a.simple_sim(1e-4, 6)            # 1 plain step + 5 graph replays
        for _ in range(3):
            b.simple_sim(1e-4, 2)        # plain launches only (host decides about the 64-bit passes)
        ga, gb = a.download(), b.download()
        assert ga.tobytes() == gb.tobytes()
        a.build_tree()
        gnodes, gidx = a.tree()
    o = parts.copy()
    orc.simple_sim(o, 1e-4, 6, order=ORDER_CANONICAL)
    scale = np.abs(o["p"]).max()
    assert np.abs(ga["p"] - o["p"]).max() / scale <= 1e-12
    onodes, oidx, _ = orc.build_tree_canonical(ga, threads=4)
    assert_tree_bit_exact(orc, gnodes, gidx, onodes, oidx, 8)


def test_quickstat_small_test_kat_gpu():
    """quickstat.rs:191-197 (the reference's one known-answer test) through kdnb_quickstat_index."""
    vals = np.array([2.3, 9.8, 3.1, 1.6, 6.7, 7.8, 8.6])
    idx = np.arange(7, dtype=np.uint64)
    kd.quickstat_index(idx, 3, vals)
    assert idx[3] == 4
    assert sorted(idx.tolist()) == list(range(7))


def test_quickstat_random_and_slices_gpu(orc):
    """quickstat.rs:199-253: `<` left of goal, `>=` right of it, entries outside the slice untouched; plus, beyond the
    reference's tests, the selected element is THE goal-th order statistic (numpy) and the oracle's quick-select
    agrees on it."""
    rng = np.random.default_rng(17)
    n = 100000
    with kd.KDTreeSim() as sim:
        for trial in range(6):
            vals = rng.random(n)
            idx = np.arange(n, dtype=np.uint64)
            start = int(rng.integers(0, n // 4)) if trial else 0
            end = start + int(rng.integers(1, 3 * n // 4)) if trial else n
            goal = int(rng.integers(start, end))
            sl = idx[start:end]                       # a view: permuted in place like `&mut indices[start..end]`
            tmp = np.ascontiguousarray(sl)
            kd.quickstat_index(tmp, goal - start, vals, sim=sim)
            sl[:] = tmp
            pivot = vals[idx[goal]]
            assert np.all(vals[idx[start:goal]] < pivot) and np.all(vals[idx[goal:end]] >= pivot)
            assert np.array_equal(idx[:start], np.arange(start, dtype=np.uint64))
            assert np.array_equal(idx[end:], np.arange(end, n, dtype=np.uint64))
            assert np.array_equal(np.sort(idx[start:end]), np.arange(start, end, dtype=np.uint64))
            assert pivot == np.partition(vals[start:end], goal - start)[goal - start]
            oidx = np.arange(start, end, dtype=np.uint64)
            orc.quickstat_index(oidx, goal - start, vals, seed=trial + 1)
            assert vals[oidx[goal - start]] == pivot


def test_quickstat_duplicates_signs_and_errors_gpu():
    """Ties, negative values and -0.0: the three blocks keep the input order (canonical permutation); panics of the
    reference (goal out of range, index out of range) become error codes."""
    rng = np.random.default_rng(5)
    vals = np.round(rng.normal(size=50000) * 4.0) / 4.0          # heavy duplicates, both signs
    vals[::97] = -0.0
    idx = rng.permutation(len(vals)).astype(np.uint64)
    before = idx.copy()
    goal = 23456
    kd.quickstat_index(idx, goal, vals)
    pivot = vals[idx[goal]]
    assert pivot == np.sort(vals)[goal]
    less, equal, greater = before[vals[before] < pivot], before[vals[before] == pivot], before[vals[before] > pivot]
    assert np.array_equal(idx, np.concatenate([less, equal, greater]))       # stable three-way partition
    with pytest.raises(kd.KdnbError):
        kd.quickstat_index(np.arange(10, dtype=np.uint64), 10, np.zeros(10))
    with pytest.raises(kd.KdnbError):
        kd.quickstat_index(np.array([0, 1, 99], dtype=np.uint64), 1, np.zeros(10))


def test_kdtree_sim_cli(orc):
    """The C++ driver mirrors Parallel/RustVersion/src/main.rs: --number/-n required, --steps/-s default 1, prints seconds."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(__file__)), "multilanguagekdtree_b200", "kdtree-sim")
    if not os.path.exists(exe):
        pytest.skip("kdtree-sim not built")
    r = subprocess.run([exe, "--number", "20000", "--steps", "3"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert float(r.stdout.strip()) > 0.0
    r = subprocess.run([exe, "-n", "500", "-s", "2", "--verbose"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "walk=" in r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=30)
    assert r.returncode == 2 and "--number" in r.stderr
    # positional form of the reference's other CLIs (Parallel/CppVersion/kdtree-sim.cpp:14-16): steps n threads
    r = subprocess.run([exe, "3", "20000", "8"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and float(r.stdout.strip()) > 0.0, r.stderr
    r = subprocess.run([exe, "3"], capture_output=True, text=True, timeout=30)
    assert r.returncode == 1
