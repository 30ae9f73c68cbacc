//! Selection (reference: Parallel/RustVersion/src/quickstat.rs).
//!
//! The reference's `quickstat_index` takes an arbitrary `lt` closure and runs on the calling thread.  Every call site
//! in the reference compares values of one array (`particles[i].p[split_dim]`, array_kd_tree.rs:561-562;
//! `vals[i]`, bin/bench_quickstat.rs:20), and that instance is what the device implements.  Inside the GPU tree build
//! the per-node quick-select does not exist any more (sorted per-dimension lists replace it), so this entry point
//! serves stand-alone callers and the benchmark.
use crate::gpu::{Context, Layout};

/// Permutes `indices` so that `indices[goal]` refers to the goal-th smallest of `vals[indices[..]]`, nothing before it
/// is larger and nothing after it is smaller — the post-condition the reference's tests check (quickstat.rs:199-253).
/// Panics where the reference panics (goal or an index out of range), and when no CUDA device is present.
pub fn quickstat_index_vals(indices: &mut [usize], goal: usize, vals: &[f64]) {
    let mut ctx = Context::new(0, 8, 0.3, Layout::Padded, 0).unwrap_or_else(|e| panic!("{}", e));
    quickstat_index_vals_on(&mut ctx, indices, goal, vals);
}

/// The same on an existing context (returns the device milliseconds of the selection).
pub fn quickstat_index_vals_on(ctx: &mut Context, indices: &mut [usize], goal: usize, vals: &[f64]) -> f64 {
    // usize is 64 bits on every platform libkdnb runs on; the compile-time check makes the cast below sound
    const _: () = assert!(std::mem::size_of::<usize>() == std::mem::size_of::<u64>());
    // SAFETY: same size and alignment, every bit pattern valid for both.
    let as_u64: &mut [u64] = unsafe { std::slice::from_raw_parts_mut(indices.as_mut_ptr() as *mut u64, indices.len()) };
    ctx.quickstat_index(vals, as_u64, goal).unwrap_or_else(|e| panic!("{}", e))
}

#[cfg(test)]
mod tests {
    use super::*;

    /// The reference's known-answer test (quickstat.rs:191-197): the 4th smallest of these values is element 4.
    #[test]
    fn small_test() {
        let vals = vec![2.3, 9.8, 3.1, 1.6, 6.7, 7.8, 8.6];
        let mut indices: Vec<usize> = (0..vals.len()).collect();
        quickstat_index_vals(&mut indices, 3, &vals);
        assert_eq!(indices[3], 4);
        for i in 0..3 {
            assert!(vals[indices[i]] <= vals[indices[3]]);
        }
        for i in 4..vals.len() {
            assert!(vals[indices[i]] >= vals[indices[3]]);
        }
    }
}
