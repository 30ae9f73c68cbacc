//! `rust_kdtree_nbody` with the step on the GPU.  Same modules as the reference crate
//! (Parallel/RustVersion/src/lib.rs:1-3) plus `gpu`, the safe wrapper around the C ABI of libkdnb.so.
//! SOURCE ONLY: never compiled in the build image (no rustc/cargo there).
pub mod array_kd_tree;
pub mod array_particle;
pub mod gpu;
pub mod quickstat;
pub mod simd_kd_tree;
pub mod simd_particle;
