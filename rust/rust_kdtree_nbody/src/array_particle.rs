//! Particle record and initial conditions (reference: Parallel/RustVersion/src/array_particle.rs).

/// Same fields, same order as the reference's `Particle` (array_particle.rs:3-8); `#[repr(C)]` so that a
/// `Vec<Particle>` is passed to libkdnb as it is (64 bytes per record = `kdnb_particle`).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct Particle {
    pub p: [f64; 3],
    pub v: [f64; 3],
    pub r: f64,
    pub m: f64,
}

/// The two-body fixture (array_particle.rs:10-17): unit mass at the origin and a test mass on the unit circle.
pub fn two_bodies() -> Vec<Particle> {
    vec![
        Particle { p: [0.0; 3], v: [0.0; 3], r: 1.0, m: 1.0 },
        Particle { p: [1.0, 0.0, 0.0], v: [0.0, 1.0, 0.0], r: 1e-4, m: 1e-20 },
    ]
}

/// splitmix64: the angle stream of `circular_orbits`.  The reference draws `fastrand::f64()` from an unseeded
/// thread-local generator (array_particle.rs:31), so no particular stream is part of its behaviour; a seeded one
/// makes runs repeatable (KDNB_SEED, default 12345 — the seed the Python mirror and the oracle use).
struct SplitMix64(u64);

impl SplitMix64 {
    fn next_f64(&mut self) -> f64 {
        self.0 = self.0.wrapping_add(0x9E37_79B9_7F4A_7C15);
        let mut z = self.0;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
        z ^= z >> 31;
        (z >> 11) as f64 * (1.0 / 9_007_199_254_740_992.0) // 53 bits -> [0, 1)
    }
}

/// Ring initial conditions (array_particle.rs:19-44): returns **n + 1** particles — the central body
/// `{p: 0, v: 0, r: 0.00465047, m: 1}` and n bodies on circular orbits with `d = 0.1 + i * 5 / n`, `v = sqrt(1 / d)`,
/// angle `u * 6.28`, `m = 1e-14`, `r = 1e-7`, all in the z = 0 plane.
pub fn circular_orbits(n: usize) -> Vec<Particle> {
    let seed = std::env::var("KDNB_SEED").ok().and_then(|s| s.parse::<u64>().ok()).unwrap_or(12345);
    let mut rng = SplitMix64(seed);
    let mut bodies = Vec::with_capacity(n + 1);
    bodies.push(Particle { p: [0.0; 3], v: [0.0; 3], r: 0.00465047, m: 1.0 });
    for i in 0..n {
        let d = 0.1 + (i as f64 * 5.0 / n as f64);
        let v = f64::sqrt(1.0 / d);
        let theta = rng.next_f64() * 6.28;
        bodies.push(Particle {
            p: [d * f64::cos(theta), d * f64::sin(theta), 0.0],
            v: [-v * f64::sin(theta), v * f64::cos(theta), 0.0],
            m: 1e-14,
            r: 1e-7,
        });
    }
    bodies
}

/// The pair acceleration the leaves of the walk apply (array_particle.rs:67-76), kept for callers that use it on
/// single pairs (tests, fixtures).  The walk itself runs on the device.
pub fn calc_pp_accel(pi: &Particle, pj: &Particle) -> [f64; 3] {
    let d = [pi.p[0] - pj.p[0], pi.p[1] - pj.p[1], pi.p[2] - pj.p[2]];
    let dist = f64::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    let magi = -pj.m / (dist * dist * dist);
    [d[0] * magi, d[1] * magi, d[2] * magi]
}
