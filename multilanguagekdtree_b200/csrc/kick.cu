// kick.cu — the kick/drift closure of simple_sim (Parallel/RustVersion/src/array_kd_tree.rs:649-662) and the
// AoS <-> SoA conversions at the C-ABI boundary (Particle, array_particle.rs:3-8).
// Bandwidth-bound elementwise kernels; arithmetic is unfused (__dmul_rn / __dadd_rn) so that results are
// bit-identical to the reference's `v += dt*a; p += dt*v` given the same accelerations.
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

// thread i = particle i (original order); its acceleration lives at tree slot rank[i]
__global__ void __launch_bounds__(256) kick_drift_kernel(uint32_t n, double dt, V3 pos, V3 vel,
                                                         const uint32_t* __restrict__ rank, AccSel sel,
                                                         PosM* __restrict__ pm) {
  pdl_sync();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t j = rank[i];
  double* acc_t = acc_buf(sel);
  const double a0 = acc_t[3 * j + 0], a1 = acc_t[3 * j + 1], a2 = acc_t[3 * j + 2];
  const double v0 = __dadd_rn(vel.p[0][i], __dmul_rn(dt, a0));  // b.v[k] += dt * a[k]   (:650-652)
  const double v1 = __dadd_rn(vel.p[1][i], __dmul_rn(dt, a1));
  const double v2 = __dadd_rn(vel.p[2][i], __dmul_rn(dt, a2));
  vel.p[0][i] = v0;
  vel.p[1][i] = v1;
  vel.p[2][i] = v2;
  const double x = __dadd_rn(pos.p[0][i], __dmul_rn(dt, v0));  // dx = dt*v; p += dx     (:653-658)
  const double y = __dadd_rn(pos.p[1][i], __dmul_rn(dt, v1));
  const double z = __dadd_rn(pos.p[2][i], __dmul_rn(dt, v2));
  pos.p[0][i] = x;
  pos.p[1][i] = y;
  pos.p[2][i] = z;
  pm[i].x = x;  // AoS mirror read by the next build
  pm[i].y = y;
  pm[i].z = z;
}

// a[k] = 0 (:659-661), coalesced over the buffer the kick just consumed
__global__ void __launch_bounds__(256) zero_acc_kernel(AccSel sel, uint64_t count) {
  pdl_sync();
  double* acc_t = acc_buf(sel);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
    acc_t[i] = 0.0;
}

int kick_drift(Ctx* c, double dt) {
  const uint32_t n = (uint32_t)c->n;
  V3 pos = {{c->pos[0], c->pos[1], c->pos[2]}}, vel = {{c->vel[0], c->vel[1], c->vel[2]}};
  KDNB_LAUNCH(c, kick_drift_kernel, (n + 255) / 256, 256, 0, n, dt, pos, vel, c->rank, acc_sel(c), c->pm);
  if (c->p2p_on) {
    KDNB_LAUNCH(c, zero_acc_kernel, 1184, 256, 0, acc_sel(c), 3ull * c->n);
  } else {
    KDNB_CUDA_TRY(c, cudaMemsetAsync(c->acc_t, 0, 3ull * c->n * sizeof(double), c->stream));
  }
  KDNB_CHECK_LAUNCH(c);
  c->tree_valid = false;  // positions moved: the tree no longer describes them
  return 0;
}

__global__ void __launch_bounds__(256) aos_to_soa_kernel(uint32_t n, const kdnb_particle* __restrict__ aos, V3 pos,
                                                         V3 vel, double* __restrict__ radius,
                                                         double* __restrict__ mass, PosM* __restrict__ pm) {
  pdl_sync();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2* q = reinterpret_cast<const double2*>(aos + i);
  const double2 a = q[0], b = q[1], c2 = q[2], d = q[3];
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

int aos_to_soa(Ctx* c) {
  const uint32_t n = (uint32_t)c->n;
  V3 pos = {{c->pos[0], c->pos[1], c->pos[2]}}, vel = {{c->vel[0], c->vel[1], c->vel[2]}};
  KDNB_LAUNCH(c, aos_to_soa_kernel, (n + 255) / 256, 256, 0, n, c->aos, pos, vel, c->radius, c->mass, c->pm);
  KDNB_CHECK_LAUNCH(c);
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
