//! Raw FFI of libkdnb.so — one declaration per entry point of include/kdnb.h.
//! SOURCE ONLY: never compiled in the build image (no rustc/cargo there).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

/// `#[repr(C)]` twin of `pub struct Particle` (Parallel/RustVersion/src/array_particle.rs:3-8): same field order, 64 bytes.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct kdnb_particle {
    pub p: [c_double; 3],
    pub v: [c_double; 3],
    pub r: c_double,
    pub m: c_double,
}

/// `#[repr(C)]` twin of the Sequential crate's SIMD particle, `Particle { p: f64x4, v: f64x4, r, m }`
/// (Sequential/RustVersion/src/simd_particle.rs:3-8): f64x4 is 32-byte aligned, 96 bytes; lane 3 of p / v is padding (0).
#[repr(C, align(32))]
#[derive(Clone, Copy, Debug, Default)]
pub struct kdnb_particle_simd {
    pub p: [c_double; 4],
    pub v: [c_double; 4],
    pub r: c_double,
    pub m: c_double,
}

pub const KDNB_LEAF: u32 = 0;
pub const KDNB_INTERNAL: u32 = 1;
pub const KDNB_NO_INDEX: u64 = u64::MAX;

/// Flat twin of `pub enum KDTree` (array_kd_tree.rs:18-34).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct kdnb_node {
    pub kind: u32,
    pub split_dim: u32,
    pub num_parts: u64,
    pub leaf_first: u64,
    pub split_val: c_double,
    pub m: c_double,
    pub cm: [c_double; 3],
    pub size: c_double,
    pub left: u64,
    pub right: u64,
}

pub const KDNB_LAYOUT_PADDED: i32 = 0;
pub const KDNB_LAYOUT_DENSE: i32 = 1;
pub const KDNB_FLAG_PROFILE: u32 = 1;
pub const KDNB_FLAG_WALK_COUNTS: u32 = 2;
pub const KDNB_FLAG_EXACT_MATH: u32 = 4;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct kdnb_config {
    pub struct_size: u32,
    pub device: i32,
    pub max_parts: u32,
    pub layout: i32,
    pub theta: c_double,
    pub flags: u32,
    pub reserved: u32,
}

#[repr(C)]
pub struct kdnb_ctx {
    _private: [u8; 0],
}

extern "C" {
    pub fn kdnb_create(cfg: *const kdnb_config) -> *mut kdnb_ctx;
    pub fn kdnb_destroy(ctx: *mut kdnb_ctx);
    pub fn kdnb_last_error(ctx: *const kdnb_ctx) -> *const c_char;
    pub fn kdnb_version() -> c_int;
    pub fn kdnb_upload_particles(ctx: *mut kdnb_ctx, aos: *const kdnb_particle, count: u64) -> c_int;
    pub fn kdnb_download_particles(ctx: *mut kdnb_ctx, out: *mut kdnb_particle, capacity: u64) -> c_int;
    pub fn kdnb_particle_count(ctx: *const kdnb_ctx) -> u64;
    pub fn kdnb_build_tree(ctx: *mut kdnb_ctx) -> c_int;
    pub fn kdnb_calc_accel(ctx: *mut kdnb_ctx) -> c_int;
    pub fn kdnb_kick_drift(ctx: *mut kdnb_ctx, dt: c_double) -> c_int;
    pub fn kdnb_simple_sim(ctx: *mut kdnb_ctx, dt: c_double, steps: i64) -> c_int;
    pub fn kdnb_simple_sim_host(cfg: *const kdnb_config, bodies: *mut kdnb_particle, count: u64, dt: c_double, steps: i64) -> c_int;
    pub fn kdnb_simple_sim_bodies(ctx: *mut kdnb_ctx, bodies: *mut kdnb_particle, count: u64, dt: c_double, steps: i64) -> c_int;
    pub fn kdnb_synchronize(ctx: *mut kdnb_ctx) -> c_int;
    pub fn kdnb_download_accel(ctx: *mut kdnb_ctx, acc: *mut c_double) -> c_int;
    pub fn kdnb_upload_accel(ctx: *mut kdnb_ctx, acc: *const c_double) -> c_int;
    pub fn kdnb_download_tree(ctx: *mut kdnb_ctx, nodes: *mut kdnb_node, cap: u64, n_nodes: *mut u64, indices: *mut u64) -> c_int;
    pub fn kdnb_download_walk_counts(ctx: *mut kdnb_ctx, counts: *mut u64) -> c_int;
    pub fn kdnb_nodes_needed(num_parts: u64, max_parts: u32) -> u64;
    pub fn kdnb_node_count(ctx: *const kdnb_ctx) -> u64;
    pub fn kdnb_shard_range(count: u64, rank: c_int, world_size: c_int, begin: *mut u64, end: *mut u64) -> c_int;
    pub fn kdnb_host_shard_range(total: u64, rank: c_int, world_size: c_int, first: *mut u64, count: *mut u64) -> c_int;
    pub fn kdnb_upload_particles_sharded(ctx: *mut kdnb_ctx, shard: *const kdnb_particle, total: u64) -> c_int;
    pub fn kdnb_download_particles_sharded(ctx: *mut kdnb_ctx, shard_out: *mut kdnb_particle) -> c_int;
    pub fn kdnb_simple_sim_bodies_sharded(ctx: *mut kdnb_ctx, shard: *mut kdnb_particle, total: u64, dt: c_double, steps: i64) -> c_int;
    pub fn kdnb_build_shard_plan(count: u64, max_parts: u32, layout: c_int, rank: c_int, world_size: c_int, first_slot: *mut u64, slots: *mut u64, first_node: *mut u64, nodes: *mut u64) -> c_int;
    pub fn kdnb_comm_unique_id(id_out_128_bytes: *mut c_void) -> c_int;
    pub fn kdnb_comm_init(ctx: *mut kdnb_ctx, id_128_bytes: *const c_void, rank: c_int, world_size: c_int) -> c_int;
    pub fn kdnb_stage_ms(ctx: *mut kdnb_ctx, ms_out: *mut c_double, steps_out: *mut u64) -> c_int;
    pub fn kdnb_stage_reset(ctx: *mut kdnb_ctx) -> c_int;
    pub fn kdnb_launch_count(ctx: *const kdnb_ctx) -> u64;
    pub fn kdnb_quickstat_index(ctx: *mut kdnb_ctx, vals: *const c_double, n_vals: u64, indices: *mut u64, count: u64, goal: u64, device_ms: *mut c_double) -> c_int;
    pub fn kdnb_measure_fp64_peak(ctx: *mut kdnb_ctx, tflops_out: *mut c_double) -> c_int;
    pub fn kdnb_flush_l2(ctx: *mut kdnb_ctx) -> c_int;
    pub fn kdnb_device_ms(ctx: *mut kdnb_ctx, begin_or_end: c_int, ms_out: *mut c_double) -> c_int;
    pub fn kdnb_upload_particles_simd(ctx: *mut kdnb_ctx, aos: *const kdnb_particle_simd, count: u64) -> c_int;
    pub fn kdnb_download_particles_simd(ctx: *mut kdnb_ctx, out: *mut kdnb_particle_simd, capacity: u64) -> c_int;
    pub fn kdnb_simple_sim_bodies_simd(ctx: *mut kdnb_ctx, bodies: *mut kdnb_particle_simd, count: u64, dt: c_double, steps: i64) -> c_int;
    pub fn kdnb_host_alloc(bytes: u64) -> *mut c_void;
    pub fn kdnb_host_free(p: *mut c_void);
}
