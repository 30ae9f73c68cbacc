//! Safe wrapper around the C ABI of libkdnb.so (include/kdnb.h, raw declarations in `kdnb-sys`).
//!
//! One `Context` owns what the reference's `simple_sim` owns for the duration of a run — `acc`, `tree`, `indices`
//! (array_kd_tree.rs:624-630) — as device buffers that are reused across calls.  There is no CPU fallback: without a
//! CUDA device `Context::new` returns an error.
use std::ffi::CStr;
use std::fmt;

use kdnb_sys as sys;

use crate::array_particle::Particle;

/// Node layout of the tree the build produces.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Layout {
    /// `build_tree_par4` (array_kd_tree.rs:515-583): right child at `cur + 1 + nodes_needed(left_len)`.
    Padded,
    /// `build_tree` (array_kd_tree.rs:63-130): right child right after the left subtree's last node.
    Dense,
}

#[derive(Debug)]
pub struct Error {
    pub code: i32,
    pub message: String,
}

impl fmt::Display for Error {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "libkdnb error {}: {}", self.code, self.message)
    }
}

impl std::error::Error for Error {}

pub type Result<T> = std::result::Result<T, Error>;

/// The flat node record of the C ABI with its tree-ordered index array.
pub struct FlatTree {
    pub nodes: Vec<sys::kdnb_node>,
    pub indices: Vec<u64>,
}

pub struct Context {
    raw: *mut sys::kdnb_ctx,
}

// One context drives one GPU from one host thread at a time (include/kdnb.h); moving it between threads is fine.
unsafe impl Send for Context {}

fn last_error(raw: *const sys::kdnb_ctx) -> String {
    // SAFETY: kdnb_last_error returns a NUL-terminated string owned by the library (or by the context).
    unsafe {
        let p = sys::kdnb_last_error(raw);
        if p.is_null() {
            String::new()
        } else {
            CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}

impl Context {
    pub fn new(device: i32, max_parts: usize, theta: f64, layout: Layout, flags: u32) -> Result<Context> {
        let cfg = sys::kdnb_config {
            struct_size: std::mem::size_of::<sys::kdnb_config>() as u32,
            device,
            max_parts: max_parts as u32,
            layout: match layout {
                Layout::Padded => sys::KDNB_LAYOUT_PADDED,
                Layout::Dense => sys::KDNB_LAYOUT_DENSE,
            },
            theta,
            flags,
            reserved: 0,
        };
        // SAFETY: cfg outlives the call; the library copies it.
        let raw = unsafe { sys::kdnb_create(&cfg) };
        if raw.is_null() {
            return Err(Error { code: -2, message: last_error(std::ptr::null()) });
        }
        Ok(Context { raw })
    }

    fn check(&self, code: i32) -> Result<()> {
        if code == 0 {
            Ok(())
        } else {
            Err(Error { code, message: last_error(self.raw) })
        }
    }

    /// The Sequential crate's SIMD surface: `simple_sim` on 96-byte f64x4 records (simd_kd_tree.rs:169-202).
    pub fn simple_sim_bodies_simd(&mut self, bodies: &mut [sys::kdnb_particle_simd], dt: f64, steps: i64) -> Result<()> {
        let rc = unsafe { sys::kdnb_simple_sim_bodies_simd(self.raw, bodies.as_mut_ptr(), bodies.len() as u64, dt, steps) };
        self.check(rc)
    }

    pub fn particle_count(&self) -> usize {
        unsafe { sys::kdnb_particle_count(self.raw) as usize }
    }

    pub fn node_count(&self) -> usize {
        unsafe { sys::kdnb_node_count(self.raw) as usize }
    }

    /// `bodies` -> device (AoS in, SoA on the device).  `Particle` is `#[repr(C)]` with the fields of `kdnb_particle`.
    pub fn upload(&mut self, bodies: &[Particle]) -> Result<()> {
        let rc = unsafe {
            sys::kdnb_upload_particles(self.raw, bodies.as_ptr() as *const sys::kdnb_particle, bodies.len() as u64)
        };
        self.check(rc)
    }

    /// Device -> `bodies`, original particle order (the reference never reorders `bodies` either).
    pub fn download(&mut self, bodies: &mut [Particle]) -> Result<()> {
        let rc = unsafe {
            sys::kdnb_download_particles(self.raw, bodies.as_mut_ptr() as *mut sys::kdnb_particle, bodies.len() as u64)
        };
        self.check(rc)
    }

    /// `indices[i] = i` + `build_tree_par4` / `build_tree` (array_kd_tree.rs:641-643).
    pub fn build_tree(&mut self) -> Result<()> {
        let rc = unsafe { sys::kdnb_build_tree(self.raw) };
        self.check(rc)
    }

    /// `acc[i] = calc_accel(i, bodies, tree)` for every particle (array_kd_tree.rs:647).
    pub fn calc_accel(&mut self) -> Result<()> {
        let rc = unsafe { sys::kdnb_calc_accel(self.raw) };
        self.check(rc)
    }

    /// `v += dt * a; p += dt * v; a = 0` (array_kd_tree.rs:649-662).
    pub fn kick_drift(&mut self, dt: f64) -> Result<()> {
        let rc = unsafe { sys::kdnb_kick_drift(self.raw, dt) };
        self.check(rc)
    }

    /// `steps` full steps on the uploaded state; asynchronous until the next download / `synchronize`.
    pub fn simple_sim(&mut self, dt: f64, steps: i64) -> Result<()> {
        let rc = unsafe { sys::kdnb_simple_sim(self.raw, dt, steps) };
        self.check(rc)
    }

    /// Upload, `steps` steps, download into `bodies` — the body of the reference's `simple_sim`.
    pub fn simple_sim_bodies(&mut self, bodies: &mut [Particle], dt: f64, steps: i64) -> Result<()> {
        let rc = unsafe {
            sys::kdnb_simple_sim_bodies(
                self.raw,
                bodies.as_mut_ptr() as *mut sys::kdnb_particle,
                bodies.len() as u64,
                dt,
                steps,
            )
        };
        self.check(rc)
    }

    pub fn synchronize(&mut self) -> Result<()> {
        let rc = unsafe { sys::kdnb_synchronize(self.raw) };
        self.check(rc)
    }

    /// The `acc` vector of the reference, original particle order.
    pub fn accel(&mut self) -> Result<Vec<[f64; 3]>> {
        let mut acc = vec![[0.0f64; 3]; self.particle_count()];
        let rc = unsafe { sys::kdnb_download_accel(self.raw, acc.as_mut_ptr() as *mut f64) };
        self.check(rc)?;
        Ok(acc)
    }

    /// `tree` and `indices` after the build, node positions exactly the reference layout's.
    pub fn tree(&mut self) -> Result<FlatTree> {
        let n_nodes = self.node_count();
        let unused = sys::kdnb_node {
            kind: sys::KDNB_LEAF,
            split_dim: 0,
            num_parts: 0,
            leaf_first: sys::KDNB_NO_INDEX,
            split_val: 0.0,
            m: 0.0,
            cm: [0.0; 3],
            size: 0.0,
            left: 0,
            right: 0,
        };
        let mut nodes = vec![unused; n_nodes];
        let mut indices = vec![0u64; self.particle_count()];
        let mut got = 0u64;
        let rc = unsafe {
            sys::kdnb_download_tree(self.raw, nodes.as_mut_ptr(), n_nodes as u64, &mut got, indices.as_mut_ptr())
        };
        self.check(rc)?;
        debug_assert_eq!(got as usize, n_nodes);
        Ok(FlatTree { nodes, indices })
    }

    /// `quickstat_index(indices, goal, |a, b| vals[a] < vals[b])` on the device (quickstat.rs:9-34).
    /// Returns the device time of the selection in milliseconds (without the copies).
    pub fn quickstat_index(&mut self, vals: &[f64], indices: &mut [u64], goal: usize) -> Result<f64> {
        let mut ms = 0.0f64;
        let rc = unsafe {
            sys::kdnb_quickstat_index(
                self.raw,
                vals.as_ptr(),
                vals.len() as u64,
                indices.as_mut_ptr(),
                indices.len() as u64,
                goal as u64,
                &mut ms,
            )
        };
        self.check(rc)?;
        Ok(ms)
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        // SAFETY: raw came from kdnb_create and is destroyed exactly once.
        unsafe { sys::kdnb_destroy(self.raw) }
    }
}
