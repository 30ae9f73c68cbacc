// common.cuh — shared device/host definitions of libkdnb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/kdnb.h"

namespace kdnb {

// ---------------------------------------------------------------------------------------------
// Device tree node used by the walk AND the download: 64 bytes = two 32-byte sectors.
//   sector 0: cx, cy, cz, m           (monopole; read only after the node is known to be internal)
//   sector 1: size2, (a, b), size, split_val
// internal: a = right child index (left = self+1), b = WN_INTERNAL | split_dim
// leaf    : a = first tree-order slot,              b = num_parts
// unused  : b = WN_UNUSED (the reference's default Leaf{0, NEGS}, array_kd_tree.rs:58)
// ---------------------------------------------------------------------------------------------
struct __align__(32) WNode {
  double cx, cy, cz, m;
  double size2;
  uint32_t a, b;
  double size;
  double split_val;
};
static_assert(sizeof(WNode) == 64, "WNode must be 64 bytes");

constexpr uint32_t WN_INTERNAL = 0x80000000u;
constexpr uint32_t WN_UNUSED = 0x40000000u;

// tree-ordered particle record read by the walk: {x, y, z, m}, 32 bytes
struct __align__(32) PosM {
  double x, y, z, m;
};

#ifndef KDNB_BOT_CAP
#define KDNB_BOT_CAP 2048  // (compile-time experiment knob: -DKDNB_BOT_CAP=1024 -DKDNB_BOT_THREADS=256, see build.py)
#endif
constexpr int BOT_CAP = KDNB_BOT_CAP;  // largest segment handled entirely in shared memory by the bottom build kernel
constexpr int MAX_LEVELS = 40;

// per-size node-count tables are closed-form; see subtree_nodes() in build.cu

struct Ctx;  // kdnb_api.cu

// ---- error plumbing
struct Status {
  int code = 0;
  std::string msg;
};

#define KDNB_CUDA_TRY(ctx, expr)                                                                    \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      return (ctx)->fail(KDNB_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    }                                                                                               \
  } while (0)

// ---- MUFU.RSQ64H: the hardware's reciprocal-square-root estimate of an f64 (relative error < 2^-22, low word 0).
// (KDNB_SIMT: the CPU execution model of tests/devtools/simt, a development aid — never defined in the product build.)
__device__ __forceinline__ double rsqrt_estimate(double x) {
#ifdef KDNB_SIMT
  return simt::rsqrt_approx(x);
#else
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
#endif
}

// ---- order-preserving key of an f64 coordinate (canonical order: -0.0 == +0.0, ties by index)
__host__ __device__ inline uint64_t f64_key(double x) {
  x = x + 0.0;  // -0.0 -> +0.0 (round-to-nearest)
#ifdef __CUDA_ARCH__
  uint64_t u = (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  __builtin_memcpy(&u, &x, 8);
#endif
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

__host__ __device__ inline uint32_t ceil_log2_u64(uint64_t k) {  // smallest e with 2^e >= k (k >= 1)
  uint32_t e = 0;
  while ((1ull << e) < k) ++e;
  return e;
}

// nodes_needed_for_particles, array_kd_tree.rs:45-53
__host__ __device__ inline uint64_t nodes_needed_padded(uint64_t n, uint32_t mp) {
  if (n <= mp) return 1;
  uint64_t k = n / (mp / 2);
  return 2 * (1ull << ceil_log2_u64(k)) - 1;
}

// number of nodes build_tree (array_kd_tree.rs:63-130) consumes for `len` particles, closed form:
// at depth j all segment sizes are floor/ceil(len / 2^j); find the first depth where all are leaves.
__host__ __device__ inline uint64_t nodes_needed_dense(uint64_t len, uint32_t mp) {
  if (len <= mp) return 1;
  uint32_t j = 1;
  while (((len + (1ull << j) - 1) >> j) > mp) ++j;  // smallest j with ceil(len/2^j) <= mp
  uint64_t half = 1ull << (j - 1);
  uint64_t q = len >> (j - 1), r = len & (half - 1);  // depth j-1: (half - r) nodes of size q, r of size q+1
  uint64_t leaves = (q <= mp) ? (half - r) + 2 * r : 2 * half;
  return 2 * leaves - 1;
}

__host__ __device__ inline uint64_t subtree_nodes(uint64_t len, uint32_t mp, int layout) {
  return layout == KDNB_LAYOUT_PADDED ? nodes_needed_padded(len, mp) : nodes_needed_dense(len, mp);
}

}  // namespace kdnb
