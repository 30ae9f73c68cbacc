//! The kD-tree simulation step (reference: Parallel/RustVersion/src/array_kd_tree.rs) with the same `pub` items,
//! executed on the GPU through libkdnb.so.  Nothing in this module loops over particles on the CPU except the
//! conversions between the library's flat node records and the reference's `KDTree` enum.
use std::fs::File;
use std::io::{BufWriter, Write};

use kdnb_sys as sys;

use crate::array_particle::Particle;
use crate::gpu::{Context, FlatTree, Layout};

pub const MAX_PARTS: usize = 8; // array_kd_tree.rs:14
pub const THETA: f64 = 0.3; // array_kd_tree.rs:15
pub const NEGS: [usize; MAX_PARTS] = [usize::MAX; MAX_PARTS]; // array_kd_tree.rs:16

/// array_kd_tree.rs:18-34 — the same variants and fields.
#[derive(Clone, Copy, Debug, PartialEq)]
pub enum KDTree {
    Leaf {
        num_parts: usize,
        leaf_parts: [usize; MAX_PARTS],
    },
    Internal {
        split_dim: usize,
        split_val: f64,
        m: f64,
        cm: [f64; 3],
        size: f64,
        left: usize,
        right: usize,
    },
}

impl KDTree {
    /// array_kd_tree.rs:37-42
    pub fn leaf(num_parts: usize, particles: [usize; MAX_PARTS]) -> KDTree {
        KDTree::Leaf { num_parts, leaf_parts: particles }
    }
}

/// array_kd_tree.rs:45-53 — `1` up to MAX_PARTS particles, else `2 * 2^ceil(log2(n / 4)) - 1` (the library evaluates
/// the closed form in integers; identical for every n, checked against the reference's f64 expression in
/// tests/test_oracle_golden.py).
pub fn nodes_needed_for_particles(num_parts: usize) -> usize {
    unsafe { sys::kdnb_nodes_needed(num_parts as u64, MAX_PARTS as u32) as usize }
}

/// array_kd_tree.rs:55-60 — every slot starts as `Leaf{0, NEGS}`.
pub fn allocate_node_vec(num_parts: usize) -> Vec<KDTree> {
    vec![KDTree::leaf(0, NEGS); nodes_needed_for_particles(num_parts)]
}

/// Flat C records -> the reference's enum.  A slot the build never wrote stays `Leaf{0, NEGS}` (:58); a written leaf
/// carries its indices followed by zeros (`[0; MAX_PARTS]`, :525).
fn store_tree(flat: &FlatTree, nodes: &mut [KDTree]) {
    for (slot, f) in nodes.iter_mut().zip(flat.nodes.iter()) {
        *slot = if f.kind == sys::KDNB_INTERNAL {
            KDTree::Internal {
                split_dim: f.split_dim as usize,
                split_val: f.split_val,
                m: f.m,
                cm: f.cm,
                size: f.size,
                left: f.left as usize,
                right: f.right as usize,
            }
        } else if f.leaf_first == sys::KDNB_NO_INDEX {
            KDTree::leaf(0, NEGS)
        } else {
            let mut parts = [0usize; MAX_PARTS];
            let first = f.leaf_first as usize;
            for k in 0..f.num_parts as usize {
                parts[k] = flat.indices[first + k] as usize;
            }
            KDTree::leaf(f.num_parts as usize, parts)
        };
    }
}

fn build_on_gpu(layout: Layout, particles: &[Particle]) -> FlatTree {
    let mut ctx = Context::new(0, MAX_PARTS, THETA, layout, 0).unwrap_or_else(|e| panic!("{}", e));
    ctx.upload(particles).unwrap_or_else(|e| panic!("{}", e));
    ctx.build_tree().unwrap_or_else(|e| panic!("{}", e));
    ctx.tree().unwrap_or_else(|e| panic!("{}", e))
}

/// array_kd_tree.rs:63-130 — the sequential build with the dense preorder node numbering; returns the last node
/// index used and grows `nodes` when it is too short (:75, :123).  The device builds whole trees: `start..end` must
/// cover all particles and `cur_node` must be 0 (every call site of the reference does that).
pub fn build_tree(
    indices: &mut Vec<usize>,
    start: usize,
    end: usize,
    particles: &Vec<Particle>,
    cur_node: usize,
    nodes: &mut Vec<KDTree>,
) -> usize {
    assert!(start == 0 && end == particles.len() && cur_node == 0, "the GPU build constructs the whole tree");
    assert!(indices.len() >= end);
    let flat = build_on_gpu(Layout::Dense, particles);
    if nodes.len() < flat.nodes.len() {
        nodes.resize(flat.nodes.len(), KDTree::leaf(0, NEGS));
    }
    store_tree(&flat, nodes);
    for (dst, src) in indices.iter_mut().zip(flat.indices.iter()) {
        *dst = *src as usize;
    }
    flat.nodes.len() - 1
}

/// array_kd_tree.rs:515-583 — the build `simple_sim` uses: padded node layout (right child at
/// `cur_node + 1 + nodes_needed_for_particles(left_len)`).  `thread_cnt` steered the rayon fork depth and has no
/// meaning on the device.
pub fn build_tree_par4(
    indices: &mut [usize],
    cur_node: usize,
    particles: &Vec<Particle>,
    nodes: &mut [KDTree],
    _thread_cnt: usize,
) {
    assert!(cur_node == 0 && indices.len() == particles.len(), "the GPU build constructs the whole tree");
    let flat = build_on_gpu(Layout::Padded, particles);
    assert!(nodes.len() >= flat.nodes.len(), "nodes is shorter than allocate_node_vec(particles.len())");
    store_tree(&flat, nodes);
    for (dst, src) in indices.iter_mut().zip(flat.indices.iter()) {
        *dst = *src as usize;
    }
}

/// `acc[i] = calc_accel(i, particles, tree)` for every i (array_kd_tree.rs:647 with :585-621) on a fresh tree of
/// `particles`.  The reference's per-particle `calc_accel(p, particles, nodes)` has no single-particle counterpart:
/// the device walks all particles of the uploaded set at once.
pub fn calc_accel_all(particles: &Vec<Particle>) -> Vec<[f64; 3]> {
    let mut ctx = Context::new(0, MAX_PARTS, THETA, Layout::Padded, 0).unwrap_or_else(|e| panic!("{}", e));
    ctx.upload(particles).unwrap_or_else(|e| panic!("{}", e));
    ctx.build_tree().unwrap_or_else(|e| panic!("{}", e));
    ctx.calc_accel().unwrap_or_else(|e| panic!("{}", e));
    ctx.accel().unwrap_or_else(|e| panic!("{}", e))
}

/// array_kd_tree.rs:623-664 — `steps` steps of build + walk + kick/drift; `bodies` is advanced in place and keeps its
/// order.  Errors panic, as the reference does.
pub fn simple_sim(bodies: &mut Vec<Particle>, dt: f64, steps: i64) {
    let mut ctx = Context::new(0, MAX_PARTS, THETA, Layout::Padded, 0).unwrap_or_else(|e| panic!("{}", e));
    ctx.simple_sim_bodies(bodies, dt, steps).unwrap_or_else(|e| panic!("{}", e));
}

/// array_kd_tree.rs:666-692 — `tree{step}.txt`: the node count, then `L n` + n lines `x y z`, or
/// `I split_dim split_val left right` (the input format of TreeVisualizer).
pub fn print_tree(step: i64, tree: &Vec<KDTree>, particles: &Vec<Particle>) -> std::io::Result<()> {
    let mut out = BufWriter::new(File::create(format!("tree{}.txt", step))?);
    writeln!(out, "{}", tree.len())?;
    for node in tree {
        match node {
            KDTree::Leaf { num_parts, leaf_parts } => {
                writeln!(out, "L {}", num_parts)?;
                for &i in &leaf_parts[..*num_parts] {
                    let p = particles[i].p;
                    writeln!(out, "{} {} {}", p[0], p[1], p[2])?;
                }
            }
            KDTree::Internal { split_dim, split_val, left, right, .. } => {
                writeln!(out, "I {} {} {} {}", split_dim, split_val, left, right)?;
            }
        }
    }
    out.flush()
}

#[cfg(test)]
mod tests {
    //! The reference's own structure tests (array_kd_tree.rs:694-878), on trees the device built.
    use super::*;
    use crate::array_particle::{circular_orbits, two_bodies};

    /// The partition invariant of array_kd_tree.rs:834-877: below an internal node every particle of the left subtree
    /// is `<= split_val` and every particle of the right subtree is `>= split_val` on `split_dim`.
    fn check_subtree(particles: &Vec<Particle>, nodes: &Vec<KDTree>, node: usize, lo: [f64; 3], hi: [f64; 3]) -> usize {
        match nodes[node] {
            KDTree::Leaf { num_parts, leaf_parts } => {
                for &i in &leaf_parts[..num_parts] {
                    for d in 0..3 {
                        assert!(particles[i].p[d] >= lo[d] && particles[i].p[d] <= hi[d], "particle {} outside its cell", i);
                    }
                }
                num_parts
            }
            KDTree::Internal { split_dim, split_val, left, right, .. } => {
                let mut hi_left = hi;
                hi_left[split_dim] = f64::min(hi[split_dim], split_val);
                let mut lo_right = lo;
                lo_right[split_dim] = f64::max(lo[split_dim], split_val);
                check_subtree(particles, nodes, left, lo, hi_left) + check_subtree(particles, nodes, right, lo_right, hi)
            }
        }
    }

    fn check_tree(particles: &Vec<Particle>, nodes: &Vec<KDTree>) {
        let n = check_subtree(particles, nodes, 0, [f64::NEG_INFINITY; 3], [f64::INFINITY; 3]);
        assert_eq!(n, particles.len());
    }

    /// array_kd_tree.rs:698-709
    #[test]
    fn single_node() {
        let parts = two_bodies();
        let mut nodes = allocate_node_vec(parts.len());
        let mut indices: Vec<usize> = (0..parts.len()).collect();
        build_tree(&mut indices, 0, parts.len(), &parts, 0, &mut nodes);
        assert!(matches!(nodes[0], KDTree::Leaf { num_parts: 2, .. }));
    }

    /// array_kd_tree.rs:711-731 — 12 particles: nodes 1 and 2 are leaves holding all of them.
    #[test]
    fn two_leaves() {
        let parts = circular_orbits(11);
        let mut nodes = allocate_node_vec(parts.len());
        let mut indices: Vec<usize> = (0..parts.len()).collect();
        build_tree(&mut indices, 0, parts.len(), &parts, 0, &mut nodes);
        match (nodes[1], nodes[2]) {
            (KDTree::Leaf { num_parts: a, .. }, KDTree::Leaf { num_parts: b, .. }) => assert_eq!(a + b, 12),
            _ => panic!("nodes 1 and 2 must be leaves"),
        }
    }

    /// array_kd_tree.rs:755-814 — 5001 particles, dense and padded builds.
    #[test]
    fn big_solar() {
        let parts = circular_orbits(5000);
        let mut indices: Vec<usize> = (0..parts.len()).collect();
        let mut nodes = allocate_node_vec(parts.len());
        build_tree(&mut indices, 0, parts.len(), &parts, 0, &mut nodes);
        check_tree(&parts, &nodes);
        let mut indices: Vec<usize> = (0..parts.len()).collect();
        let mut nodes = allocate_node_vec(parts.len());
        build_tree_par4(&mut indices, 0, &parts, &mut nodes, 1);
        check_tree(&parts, &nodes);
    }

    /// array_kd_tree.rs:816-832 — the invariant still holds on a tree built after 10 steps.
    #[test]
    fn big_solar_with_steps() {
        let mut parts = circular_orbits(5000);
        simple_sim(&mut parts, 1e-3, 10);
        let mut indices: Vec<usize> = (0..parts.len()).collect();
        let mut nodes = allocate_node_vec(parts.len());
        build_tree_par4(&mut indices, 0, &parts, &mut nodes, 1);
        check_tree(&parts, &nodes);
    }

    /// Analytic anchor: the two-body fixture with dt = pi / 1000 is back at (1, 0) after 2000 steps.
    #[test]
    fn two_bodies_full_orbit() {
        let mut parts = two_bodies();
        simple_sim(&mut parts, std::f64::consts::PI / 1000.0, 2000);
        assert!((parts[1].p[0] - 1.0).abs() < 1e-2 && parts[1].p[1].abs() < 1e-2);
    }
}
