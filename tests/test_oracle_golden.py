"""Pin the CPU oracle against the reference's own outputs (tests/golden, made by make_golden.py from the
reference's PureVersion Python) and against the reference's known-answer / structural tests.
Everything here is BIT-EXACT unless a tolerance is written next to the assertion."""
import glob
import math
import os

import numpy as np
import pytest

from oracle.okd import LAYOUT_DENSE, LAYOUT_PADDED, ORDER_CANONICAL, ORDER_FAITHFUL, PARTICLE

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLD, "pure*.npz")))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_matches_reference_python_bit_exact(orc, path):
    g = np.load(path)
    parts0 = g["parts0"].view(PARTICLE).reshape(-1)
    mp, seed = int(g["max_parts"]), int(g["pivot_seed"])
    nodes, idx, last = orc.build_tree(parts0, max_parts=mp, seed=seed)
    assert last == int(g["last"])
    assert np.array_equal(idx, g["indices"])
    t = nodes[: last + 1]
    internal = g["tree_is_internal"].astype(bool)
    assert np.array_equal(t["is_internal"].astype(bool), internal)
    leaf = ~internal
    assert np.array_equal(t["num_parts"][leaf], g["tree_num_parts"][leaf])
    for i in np.nonzero(leaf)[0]:
        k = int(g["tree_num_parts"][i])
        assert np.array_equal(t["leaf_parts"][i][:k], g["tree_leaf_parts"][i][:k])
    for f in ("split_dim", "left", "right"):
        assert np.array_equal(t[f][internal], g["tree_" + f][internal]), f
    if "pure3d" in path:  # the 3-D case exists to pin the choice of z: it must actually occur
        assert (g["tree_split_dim"][internal] == 2).sum() >= 5
    for f in ("split_val", "m", "cm", "size"):
        assert np.array_equal(bits(t[f][internal]), bits(g["tree_" + f][internal])), f
    # calc_accel: pairwise recursion, bit-exact
    acc = orc.calc_accel_all(parts0, nodes, theta=0.3)
    assert np.array_equal(bits(acc), bits(g["acc"]))
    # simple_sim trajectory (dense layout, sequential pivot stream), bit-exact
    sim = parts0.copy()
    orc.simple_sim(sim, float(g["dt"]), int(g["steps"]), max_parts=mp, theta=0.3, layout=LAYOUT_DENSE,
                   order=ORDER_FAITHFUL, seed=seed, threads=1)
    after = g["after"].view(PARTICLE).reshape(-1)
    assert np.array_equal(bits(sim["p"]), bits(after["p"]))
    assert np.array_equal(bits(sim["v"]), bits(after["v"]))


def test_two_bodies_half_orbit_analytic(orc):
    # G=M=r=1 circular orbit: after 1000 steps of dt=pi/1000 the light body is at ~(-1,0) (symplectic Euler, O(dt)).
    g = np.load(os.path.join(GOLD, "pure_two_bodies.npz"))
    after = g["after"].view(PARTICLE).reshape(-1)
    assert abs(after["p"][1][0] + 1.0) < 2e-2 and abs(after["p"][1][1]) < 2e-2


def test_quickstat_small_test_kat(orc):
    # Parallel/RustVersion/src/quickstat.rs:191-197
    vals = np.array([2.3, 9.8, 3.1, 1.6, 6.7, 7.8, 8.6])
    for seed in range(20):
        idx = np.arange(7, dtype=np.uint64)
        orc.quickstat_index(idx, 3, vals, seed=seed)
        assert idx[3] == 4


def test_quickstat_random_slices(orc):
    # quickstat.rs:199-253 (post-condition: `<` left of goal, `>=` right, outside the slice untouched)
    rng = np.random.default_rng(5)
    n = 20000
    for trial in range(20):
        vals = rng.random(n)
        idx = np.arange(n, dtype=np.uint64)
        start = int(rng.integers(0, n // 4))
        end = start + int(rng.integers(1, 3 * n // 4))
        goal = int(rng.integers(start, end))
        sl = idx[start:end]
        orc.quickstat_index(sl, goal - start, vals, seed=trial)
        assert np.array_equal(idx[:start], np.arange(start)) and np.array_equal(idx[end:], np.arange(end, n))
        pv = vals[idx[goal]]
        assert np.all(vals[idx[start:goal]] < pv) and np.all(vals[idx[goal:end]] >= pv)


def test_nodes_needed_matches_reference_formula(orc):
    # array_kd_tree.rs:45-53 evaluated in float like the reference does
    for mp in (4, 7, 8, 16):
        for n in list(range(1, 300)) + [1001, 5001, 100001, 1000001, 10000001, 100000001]:
            if n <= mp:
                want = 1
            else:
                want = 2 * 2 ** math.ceil(math.log2(float(n // (mp // 2)))) - 1
            assert orc.nodes_needed_for_particles(n, mp) == want
    assert orc.nodes_needed_for_particles(1000001, 8) == 524287
    assert orc.nodes_needed_for_particles(10000001, 8) == 8388607


def test_single_node(orc):
    # array_kd_tree.rs:698-709 (its stale `len()==2` assertion is not ported)
    parts = orc.two_bodies()
    nodes, idx, last = orc.build_tree(parts)
    assert last == 0 and nodes[0]["is_internal"] == 0 and nodes[0]["num_parts"] == 2
    assert orc.nodes_needed_for_particles(2, 8) == 1


@pytest.mark.parametrize("builder", ["dense", "par4", "canon_padded", "canon_dense"])
def test_two_leaves(orc, builder):
    # array_kd_tree.rs:711-753
    parts = orc.circular_orbits(11, seed=42)
    nodes = _build(orc, builder, parts)
    assert orc.check_tree_struct(nodes, parts) == 0
    assert nodes[0]["is_internal"] == 1
    assert nodes[1]["is_internal"] == 0 and nodes[2]["is_internal"] == 0
    assert nodes[1]["num_parts"] + nodes[2]["num_parts"] == 12


def _build(orc, builder, parts, threads=1):
    if builder == "dense":
        return orc.build_tree(parts)[0]
    if builder == "par4":
        return orc.build_tree_par4(parts, threads=threads)[0]
    if builder == "canon_padded":
        return orc.build_tree_canonical(parts, layout=LAYOUT_PADDED, threads=threads)[0]
    return orc.build_tree_canonical(parts, layout=LAYOUT_DENSE, threads=threads)[0]


@pytest.mark.parametrize("builder", ["dense", "par4", "canon_padded", "canon_dense"])
def test_big_solar(orc, builder):
    # array_kd_tree.rs:755-814
    parts = orc.circular_orbits(5000, seed=9)
    nodes = _build(orc, builder, parts, threads=4)
    assert orc.check_tree_struct(nodes, parts) == 0


def test_big_solar_with_steps(orc):
    # array_kd_tree.rs:816-832
    parts = orc.circular_orbits(5000, seed=9)
    orc.simple_sim(parts, 1e-3, 10)
    nodes, _ = orc.build_tree_par4(parts, threads=2)
    assert orc.check_tree_struct(nodes, parts) == 0


def _tree_sets(nodes):
    """order-independent signature of a tree: per node (kind, split_dim, split_val bits, size bits, left, right, sorted leaf set)"""
    sig = []
    for nd in nodes:
        if nd["is_internal"]:
            sig.append((1, int(nd["split_dim"]), float(nd["split_val"]).hex(), float(nd["size"]).hex(), int(nd["left"]), int(nd["right"])))
        else:
            k = int(nd["num_parts"])
            sig.append((0, k, tuple(sorted(int(x) for x in nd["leaf_parts"][:k]))))
    return sig


@pytest.mark.parametrize("n", [11, 100, 1000, 5000, 20000])
def test_faithful_and_canonical_agree_on_order_independent_fields(orc, n):
    """SURVEY.md §0 finding 2: different pivots change only in-leaf order and the rounding of m/cm."""
    parts = orc.circular_orbits(n, seed=n)
    a, _ = orc.build_tree_par4(parts, seed=1)
    b, _ = orc.build_tree_par4(parts, seed=2, threads=4)
    c, _, _ = orc.build_tree_canonical(parts, layout=LAYOUT_PADDED, threads=3)
    assert _tree_sets(a) == _tree_sets(b) == _tree_sets(c)
    ia = a["is_internal"].astype(bool)
    for other in (b, c):
        assert np.allclose(other["m"][ia], a["m"][ia], rtol=1e-11, atol=0)          # tolerance: summation order only (reference self-noise ~ n*eps)
        assert np.allclose(other["cm"][ia], a["cm"][ia], rtol=0, atol=1e-11)
    # dense layouts: same nodes, preorder numbering
    d, _, last_d = orc.build_tree(parts, seed=3)
    e, _, last_e = orc.build_tree_canonical(parts, layout=LAYOUT_DENSE)
    assert last_d == last_e
    assert _tree_sets(d[: last_d + 1]) == _tree_sets(e[: last_e + 1])
    # walk on faithful vs canonical tree: agreement to 1e-12 relative (stated tolerance of the path)
    acc_a = orc.calc_accel_all(parts, a)
    acc_c = orc.calc_accel_all(parts, c)
    rel = np.linalg.norm(acc_a - acc_c, axis=1) / np.linalg.norm(acc_a, axis=1)
    assert rel.max() < 1e-12


def test_canonical_leaves_ascending_and_sim_modes_agree(orc):
    parts = orc.circular_orbits(3000, seed=21)
    c, idx, _ = orc.build_tree_canonical(parts)
    leaves = c[(c["is_internal"] == 0) & (c["num_parts"] > 0)]
    for nd in leaves:
        k = int(nd["num_parts"])
        assert np.all(np.diff(nd["leaf_parts"][:k].astype(np.int64)) > 0)
    assert sorted(idx.tolist()) == list(range(len(parts)))
    s1, s2 = parts.copy(), parts.copy()
    orc.simple_sim(s1, 1e-3, 10, order=ORDER_FAITHFUL, seed=5)
    orc.simple_sim(s2, 1e-3, 10, order=ORDER_CANONICAL)
    rel = np.linalg.norm(s1["p"] - s2["p"], axis=1).max() / np.abs(s1["p"]).max()
    assert rel < 1e-12   # stated tolerance for trajectories


def test_threads_do_not_change_results(orc):
    parts = orc.circular_orbits(20000, seed=4)
    a, ia = orc.build_tree_par4(parts, seed=9, threads=1)
    b, ib = orc.build_tree_par4(parts, seed=9, threads=8)
    assert a.tobytes() == b.tobytes() and np.array_equal(ia, ib)


def test_print_tree_format(orc, tmp_path):
    # array_kd_tree.rs:666-692 / TreeVisualizer/src/main/scala/ViewTrees.scala:32-80
    parts = orc.circular_orbits(100, seed=1)
    nodes, _ = orc.build_tree_par4(parts)
    p = tmp_path / "tree0.txt"
    orc.print_tree(str(p), nodes, parts)
    lines = p.read_text().split("\n")
    assert int(lines[0]) == len(nodes)
    assert lines[1].startswith("I ")
