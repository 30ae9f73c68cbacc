//! The Sequential crate's SIMD particle (reference: Sequential/RustVersion/src/simd_particle.rs) over the C ABI.
//! SOURCE ONLY: never compiled in the build image (no rustc/cargo there).
use kdnb_sys as sys;

/// `Particle { p: f64x4, v: f64x4, r: f64, m: f64 }` (simd_particle.rs:3-8) as the FFI-safe 96-byte record
/// (`f64x4` is 32-byte aligned); lane 3 of `p` and `v` is padding and must stay 0.
pub type Particle = sys::kdnb_particle_simd;

/// simd_particle.rs:10-27
pub fn two_bodies() -> Vec<Particle> {
    vec![
        Particle { p: [0.0; 4], v: [0.0; 4], r: 1.0, m: 1.0 },
        Particle { p: [1.0, 0.0, 0.0, 0.0], v: [0.0, 1.0, 0.0, 0.0], r: 1e-4, m: 1e-20 },
    ]
}

/// simd_particle.rs:29-55: n + 1 particles; the ring angles are `u * TAU` here (the array version uses 6.28).
pub fn circular_orbits(n: usize) -> Vec<Particle> {
    let mut s: u64 = 12345;
    let mut next = move || {
        s = s.wrapping_add(0x9E37_79B9_7F4A_7C15);
        let mut z = s;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
        z ^= z >> 31;
        (z >> 11) as f64 * (1.0 / 9_007_199_254_740_992.0)
    };
    let mut buf = Vec::with_capacity(n + 1);
    buf.push(Particle { p: [0.0; 4], v: [0.0; 4], r: 0.00465047, m: 1.0 });
    for i in 0..n {
        let d = 0.1 + ((i as f64) * 5.0 / (n as f64));
        let v = f64::sqrt(1.0 / d);
        let theta = next() * std::f64::consts::TAU;
        buf.push(Particle {
            p: [d * f64::cos(theta), d * f64::sin(theta), 0.0, 0.0],
            v: [-v * f64::sin(theta), v * f64::cos(theta), 0.0, 0.0],
            r: 1e-7,
            m: 1e-14,
        });
    }
    buf
}
