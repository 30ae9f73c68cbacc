// kick.cu — the kick/drift closure of simple_sim (Parallel/RustVersion/src/array_kd_tree.rs:649-662) and the
// AoS <-> SoA conversions at the C-ABI boundary (Particle, array_particle.rs:3-8).
// Bandwidth-bound elementwise kernels; arithmetic is unfused (__dmul_rn / __dadd_rn) so that results are
// bit-identical to the reference's `v += dt*a; p += dt*v` given the same accelerations.
#include <algorithm>

#include "ctx.cuh"

namespace kdnb {

struct V3 {
  double* p[3];
};

// which acc buffer holds the accelerations of the step being consumed: peer mode alternates two buffers by step
// parity (walk wrote buffer epoch&1, the wait kernel then incremented epoch); otherwise there is one buffer
struct AccSel {
  double* base;
  const uint32_t* epoch;  // nullptr: single buffer
  uint64_t stride;
};
__device__ __forceinline__ double* acc_buf(const AccSel& a) {
  return a.epoch ? a.base + (uint64_t)((*a.epoch + 1u) & 1u) * a.stride : a.base;
}
static AccSel acc_sel(const Ctx* c) {
  AccSel a;
  a.base = c->acc_t;
  a.epoch = c->p2p_on ? c->p2p_state : nullptr;
  a.stride = c->acc_stride;
  return a;
}

// thread i = particle i (original order); its acceleration lives at tree slot rank[i].  The same thread folds the new
// position into the extents the next build's sort scales its keys to (accumulate_extent, common.cuh): the step has no
// separate pass over the positions.  `a[k] = 0` (:659-661) costs no memory traffic: the context remembers that the
// accelerations were consumed (acc_valid = false) — a kick without a walk in between then runs with use_acc = false
// (a = 0, exactly what the reference's zeroed vector gives), and downloading them yields zeros.  (Measured: zeroing the
// gathered 24 bytes in place doubles the kernel's time at N = 10M, a separate memset is a launch and 24 B / particle.)
__global__ void __launch_bounds__(256) kick_drift_kernel(uint32_t n, double dt, V3 pos, V3 vel,
                                                         const double* __restrict__ mass,
                                                         const uint32_t* __restrict__ rank, AccSel sel, bool use_acc,
                                                         PosM* __restrict__ pm, uint64_t* __restrict__ ss) {
  pdl_sync();
  __shared__ uint64_t sm[8][8];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  double x = 0.0, y = 0.0, z = 0.0;
  if (valid) {
    // every load of the particle before its first store: the arrays are not provably disjoint for the compiler, so a
    // store in between would order the later loads behind it (five dependent round trips instead of two)
    const double* acc_t = acc_buf(sel);
    const uint64_t j = rank[i];
    const double u0 = vel.p[0][i], u1 = vel.p[1][i], u2 = vel.p[2][i];
    const double p0 = pos.p[0][i], p1 = pos.p[1][i], p2 = pos.p[2][i];
    const double mi = mass[i];  // (8 bytes read so that the mirror record below is written whole: a 24-of-32-byte
                                // write makes the L2 fetch the sector first)
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (use_acc) a0 = acc_t[3 * j + 0], a1 = acc_t[3 * j + 1], a2 = acc_t[3 * j + 2];
    const double v0 = __dadd_rn(u0, __dmul_rn(dt, a0));  // b.v[k] += dt * a[k]   (:650-652)
    const double v1 = __dadd_rn(u1, __dmul_rn(dt, a1));
    const double v2 = __dadd_rn(u2, __dmul_rn(dt, a2));
    x = __dadd_rn(p0, __dmul_rn(dt, v0));  // dx = dt*v; p += dx     (:653-658)
    y = __dadd_rn(p1, __dmul_rn(dt, v1));
    z = __dadd_rn(p2, __dmul_rn(dt, v2));
    vel.p[0][i] = v0;
    vel.p[1][i] = v1;
    vel.p[2][i] = v2;
    pos.p[0][i] = x;
    pos.p[1][i] = y;
    pos.p[2][i] = z;
    PosM rec;  // AoS mirror read by the next build
    rec.x = x, rec.y = y, rec.z = z, rec.m = mi;
    pm[i] = rec;
  }
  accumulate_extent(x, y, z, valid, 2u, ss, sm);  // (2: the masses did not change, sort_prep keeps what the upload found)
}

int kick_drift(Ctx* c, double dt) {
  const uint32_t n = (uint32_t)c->n;
  V3 pos = {{c->pos[0], c->pos[1], c->pos[2]}}, vel = {{c->vel[0], c->vel[1], c->vel[2]}};
  KDNB_LAUNCH(c, kick_drift_kernel, (n + 255) / 256, 256, 0, n, dt, pos, vel, c->mass, c->rank, acc_sel(c), c->acc_valid, c->pm,
              c->sort_state);
  KDNB_CHECK_LAUNCH(c);
  c->extent_fresh = true;
  c->acc_valid = false;   // a = 0: consumed
  c->tree_valid = false;  // positions moved: the tree no longer describes them
  return 0;
}

// the extent records start empty before a new particle set is converted
__global__ void __launch_bounds__(EXT_PARTS) extent_reset_kernel(uint64_t* ss) {
  pdl_sync();
  uint64_t* rec = ss + SS_PART + 8 * threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) rec[j] = j < 3 ? ~0ull : 0ull;
}

__global__ void __launch_bounds__(256) aos_to_soa_kernel(uint32_t n, const kdnb_particle* __restrict__ aos, V3 pos,
                                                         V3 vel, double* __restrict__ radius,
                                                         double* __restrict__ mass, PosM* __restrict__ pm,
                                                         uint64_t* __restrict__ ss) {
  pdl_sync();
  __shared__ uint64_t sm[8][8];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  double x = 0.0, y = 0.0, z = 0.0;
  uint32_t light = 0u;
  if (valid) {
    const double2* q = reinterpret_cast<const double2*>(aos + i);
    const double2 a = q[0], b = q[1], c2 = q[2], d = q[3];
    x = a.x, y = a.y, z = b.x;
    pos.p[0][i] = a.x;
    pos.p[1][i] = a.y;
    pos.p[2][i] = b.x;
    vel.p[0][i] = b.y;
    vel.p[1][i] = c2.x;
    vel.p[2][i] = c2.y;
    radius[i] = d.x;
    mass[i] = d.y;
    PosM rec;
    rec.x = a.x;
    rec.y = a.y;
    rec.z = b.x;
    rec.m = d.y;
    pm[i] = rec;
    light = (d.y > 0.0) ? 0u : 1u;
  }
  accumulate_extent(x, y, z, valid, light, ss, sm);
}

__global__ void __launch_bounds__(256) soa_to_aos_kernel(uint32_t n, kdnb_particle* __restrict__ aos, V3 pos, V3 vel,
                                                         const double* __restrict__ radius,
                                                         const double* __restrict__ mass) {
  pdl_sync();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2* q = reinterpret_cast<double2*>(aos + i);
  q[0] = make_double2(pos.p[0][i], pos.p[1][i]);
  q[1] = make_double2(pos.p[2][i], vel.p[0][i]);
  q[2] = make_double2(vel.p[1][i], vel.p[2][i]);
  q[3] = make_double2(radius[i], mass[i]);
}

// the Sequential crate's SIMD particle (simd_particle.rs:3-8): 96-byte records {p[4], v[4], r, m, pad}.  Lane 3 is
// padding in the reference (always 0); a record with a non-zero lane 3 raises the context's input-error flag.
__global__ void __launch_bounds__(256) simd_to_soa_kernel(uint32_t n, const kdnb_particle_simd* __restrict__ aos, V3 pos,
                                                          V3 vel, double* __restrict__ radius,
                                                          double* __restrict__ mass, PosM* __restrict__ pm,
                                                          uint64_t* __restrict__ ss, uint32_t* __restrict__ bad) {
  pdl_sync();
  __shared__ uint64_t sm[8][8];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  double x = 0.0, y = 0.0, z = 0.0;
  uint32_t light = 0u;
  if (valid) {
    const double2* q = reinterpret_cast<const double2*>(aos + i);
    const double2 p01 = q[0], p23 = q[1], v01 = q[2], v23 = q[3], rm = q[4];
    if (p23.y != 0.0 || v23.y != 0.0) *bad = 1u;  // (also NaN: not a padding lane)
    x = p01.x, y = p01.y, z = p23.x;
    pos.p[0][i] = x;
    pos.p[1][i] = y;
    pos.p[2][i] = z;
    vel.p[0][i] = v01.x;
    vel.p[1][i] = v01.y;
    vel.p[2][i] = v23.x;
    radius[i] = rm.x;
    mass[i] = rm.y;
    PosM rec;
    rec.x = x, rec.y = y, rec.z = z, rec.m = rm.y;
    pm[i] = rec;
    light = (rm.y > 0.0) ? 0u : 1u;
  }
  accumulate_extent(x, y, z, valid, light, ss, sm);
}

__global__ void __launch_bounds__(256) soa_to_simd_kernel(uint32_t n, kdnb_particle_simd* __restrict__ aos, V3 pos, V3 vel,
                                                          const double* __restrict__ radius,
                                                          const double* __restrict__ mass) {
  pdl_sync();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2* q = reinterpret_cast<double2*>(aos + i);
  q[0] = make_double2(pos.p[0][i], pos.p[1][i]);
  q[1] = make_double2(pos.p[2][i], 0.0);  // lane 3 stays 0: v[3] += dt * 0, p[3] += dt * v[3] (simd_kd_tree.rs:195-198)
  q[2] = make_double2(vel.p[0][i], vel.p[1][i]);
  q[3] = make_double2(vel.p[2][i], 0.0);
  q[4] = make_double2(radius[i], mass[i]);
  q[5] = make_double2(0.0, 0.0);
}

int simd_to_soa(Ctx* c, const kdnb_particle_simd* dev_aos) {
  const uint32_t n = (uint32_t)c->n;
  V3 pos = {{c->pos[0], c->pos[1], c->pos[2]}}, vel = {{c->vel[0], c->vel[1], c->vel[2]}};
  KDNB_LAUNCH(c, extent_reset_kernel, 1, EXT_PARTS, 0, c->sort_state);
  KDNB_LAUNCH(c, simd_to_soa_kernel, (n + 255) / 256, 256, 0, n, dev_aos, pos, vel, c->radius, c->mass, c->pm,
              c->sort_state, c->lvl_ctl + 66);
  KDNB_CHECK_LAUNCH(c);
  c->extent_fresh = true;
  return 0;
}

int soa_to_simd(Ctx* c, kdnb_particle_simd* dev_aos) {
  const uint32_t n = (uint32_t)c->n;
  V3 pos = {{c->pos[0], c->pos[1], c->pos[2]}}, vel = {{c->vel[0], c->vel[1], c->vel[2]}};
  KDNB_LAUNCH(c, soa_to_simd_kernel, (n + 255) / 256, 256, 0, n, dev_aos, pos, vel, c->radius, c->mass);
  KDNB_CHECK_LAUNCH(c);
  return 0;
}

int aos_to_soa(Ctx* c) {
  const uint32_t n = (uint32_t)c->n;
  V3 pos = {{c->pos[0], c->pos[1], c->pos[2]}}, vel = {{c->vel[0], c->vel[1], c->vel[2]}};
  KDNB_LAUNCH(c, extent_reset_kernel, 1, EXT_PARTS, 0, c->sort_state);
  KDNB_LAUNCH(c, aos_to_soa_kernel, (n + 255) / 256, 256, 0, n, c->aos, pos, vel, c->radius, c->mass, c->pm, c->sort_state);
  KDNB_CHECK_LAUNCH(c);
  c->extent_fresh = true;
  return 0;
}

int soa_to_aos(Ctx* c) {
  const uint32_t n = (uint32_t)c->n;
  V3 pos = {{c->pos[0], c->pos[1], c->pos[2]}}, vel = {{c->vel[0], c->vel[1], c->vel[2]}};
  KDNB_LAUNCH(c, soa_to_aos_kernel, (n + 255) / 256, 256, 0, n, c->aos, pos, vel, c->radius, c->mass);
  KDNB_CHECK_LAUNCH(c);
  return 0;
}

__global__ void __launch_bounds__(256) gather_acc_kernel(uint32_t n, const uint32_t* __restrict__ rank, AccSel sel,
                                                         double* __restrict__ out) {
  pdl_sync();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t j = rank[i];
  const double* acc_t = acc_buf(sel);
  out[3ull * i + 0] = acc_t[3 * j + 0];
  out[3ull * i + 1] = acc_t[3 * j + 1];
  out[3ull * i + 2] = acc_t[3 * j + 2];
}

__global__ void __launch_bounds__(256) scatter_acc_kernel(uint32_t n, const uint32_t* __restrict__ rank,
                                                          const double* __restrict__ in, AccSel sel) {
  pdl_sync();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t j = rank[i];
  double* acc_t = acc_buf(sel);
  acc_t[3 * j + 0] = in[3ull * i + 0];
  acc_t[3 * j + 1] = in[3ull * i + 1];
  acc_t[3 * j + 2] = in[3ull * i + 2];
}

__global__ void __launch_bounds__(256) gather_counts_kernel(uint32_t n, const uint32_t* __restrict__ rank,
                                                            const unsigned long long* __restrict__ in,
                                                            unsigned long long* __restrict__ out) {
  pdl_sync();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t j = rank[i];
#pragma unroll
  for (int k = 0; k < 4; ++k) out[4ull * i + k] = in[4 * j + k];
}

int gather_acc(Ctx* c, double* dst) {
  const uint32_t n = (uint32_t)c->n;
  KDNB_LAUNCH(c, gather_acc_kernel, (n + 255) / 256, 256, 0, n, c->rank, acc_sel(c), dst);
  KDNB_CHECK_LAUNCH(c);
  return 0;
}

int scatter_acc(Ctx* c, const double* src) {
  const uint32_t n = (uint32_t)c->n;
  KDNB_LAUNCH(c, scatter_acc_kernel, (n + 255) / 256, 256, 0, n, c->rank, src, acc_sel(c));
  KDNB_CHECK_LAUNCH(c);
  return 0;
}

int gather_counts(Ctx* c, unsigned long long* dst) {
  const uint32_t n = (uint32_t)c->n;
  KDNB_LAUNCH(c, gather_counts_kernel, (n + 255) / 256, 256, 0, n, c->rank, c->wcounts, dst);
  KDNB_CHECK_LAUNCH(c);
  return 0;
}

}  // namespace kdnb
