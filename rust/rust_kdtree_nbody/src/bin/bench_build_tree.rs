//! Parallel/RustVersion/src/bin/bench_build_tree.rs:1-42 — tree build at N = 10M.  Of the reference's seven variants
//! the two whose results `simple_sim` and the tests use exist here: `build_tree` (dense) and `build_tree_par4`
//! (padded).  Timed on the device context directly so that the tree download is not part of the figure.
use std::time::Instant;

use rust_kdtree_nbody::array_kd_tree::{MAX_PARTS, THETA};
use rust_kdtree_nbody::array_particle::circular_orbits;
use rust_kdtree_nbody::gpu::{Context, Layout};

fn time_build(name: &str, layout: Layout, parts: &[rust_kdtree_nbody::array_particle::Particle]) {
    let mut ctx = Context::new(0, MAX_PARTS, THETA, layout, 0).expect("kdnb_create");
    ctx.upload(parts).expect("upload");
    ctx.build_tree().expect("warm-up build");
    ctx.synchronize().expect("synchronize");
    let pre = Instant::now();
    ctx.build_tree().expect("build");
    ctx.synchronize().expect("synchronize");
    eprintln!("{} Runtime = {}", name, pre.elapsed().as_secs_f64());
}

fn main() {
    let parts = circular_orbits(10_000_000);
    time_build("Sequential-layout (dense)", Layout::Dense, &parts);
    time_build("Par4-layout (padded)", Layout::Padded, &parts);
}
