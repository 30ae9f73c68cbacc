//! Parallel/RustVersion/src/bin/bench_quickstat.rs:1-49 — median selection over n = 100M values.
use rust_kdtree_nbody::gpu::{Context, Layout};
use rust_kdtree_nbody::quickstat::quickstat_index_vals_on;

fn main() {
    let n: usize = 100_000_000;
    // deterministic pseudo-random values in [0, 1) (the reference uses an unseeded generator)
    let mut state = 0x1234_5678_9ABC_DEF0u64;
    let vals: Vec<f64> = (0..n)
        .map(|_| {
            state ^= state << 13;
            state ^= state >> 7;
            state ^= state << 17;
            (state >> 11) as f64 / 9_007_199_254_740_992.0
        })
        .collect();
    let mut indices: Vec<usize> = (0..n).collect();
    let mut ctx = Context::new(0, 8, 0.3, Layout::Padded, 0).expect("kdnb_create");
    let ms = quickstat_index_vals_on(&mut ctx, &mut indices, n / 2, &vals);
    eprintln!("Device selection = {} s (without the host<->device copies)", ms / 1e3);
    let pivot = vals[indices[n / 2]];
    assert!(indices[..n / 2].iter().all(|&i| vals[i] <= pivot));
    assert!(indices[n / 2..].iter().all(|&i| vals[i] >= pivot));
}
