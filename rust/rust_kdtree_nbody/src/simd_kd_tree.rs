//! The Sequential crate's SIMD step (reference: Sequential/RustVersion/src/simd_kd_tree.rs:169-202) on the GPU.
//! SOURCE ONLY: never compiled in the build image (no rustc/cargo there).
use crate::gpu::{Context, Layout};
use crate::simd_particle::Particle;

pub const MAX_PARTS: usize = 7; // simd_kd_tree.rs:9
pub const THETA: f64 = 0.3; // simd_kd_tree.rs:10

/// `simple_sim(bodies, dt, steps)`: MAX_PARTS = 7, dense `build_tree` layout; bodies are advanced in place.
/// Panics where the reference would (and when no CUDA device is available: there is no CPU fallback).
pub fn simple_sim(bodies: &mut Vec<Particle>, dt: f64, steps: i64) {
    let mut ctx = Context::new(0, MAX_PARTS, THETA, Layout::Dense, 0).expect("kdnb_create");
    ctx.simple_sim_bodies_simd(bodies, dt, steps).expect("kdnb_simple_sim_bodies_simd");
}
