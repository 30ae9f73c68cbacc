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
#define KDNB_BOT_CAP 1024  // 1024 slots x 256 threads x 6 CTAs/SM: measured against 2048 x 512 x 4 (profiles/r02_ab_build_botcap.txt:
                           // build 0.514 -> 0.487 ms at N=1M, 4.45 -> 3.96 ms at N=10M); compile-time knob with KDNB_BOT_THREADS / _MINB
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

// ---- look-back status words (build.cu): single 64-bit volatile accesses
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
#ifdef KDNB_SIMT
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
#endif
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
#ifdef KDNB_SIMT
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}
// ---- hide a value's provenance from the compiler (it stays in its registers; nothing is emitted)
#ifdef KDNB_SIMT
#define KDNB_OPAQUE_F64(x) ((void)0)
#else
#define KDNB_OPAQUE_F64(x) asm volatile("" : "+d"(x))
#endif
// ---- base of the dynamic shared memory of the running CTA
#ifdef KDNB_SIMT
#define KDNB_DYN_SMEM(name) unsigned char* const name = simt::dyn_smem()
#else
#define KDNB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

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

// sort state (device, sort.cu): [0..2] lo, [3..5] scale of the 32-bit keys as f64 bits (sort_prep), [6] need64,
// [7] "some mass is <= 0" (set at upload); from word SS_PART on, EXT_PARTS records of 8 words: running extreme keys
// {min x, y, z, max x, y, z, light (0 / 1; 2 = the writer did not see the masses), -} of the positions, accumulated by
// the kernel that last WROTE them (aos_to_soa, kick_drift), consumed and reset by sort_prep at the start of the next build
constexpr int SS_LO = 0, SS_SCALE = 3, SS_NEED64 = 6, SS_LIGHT = 7, SS_PART = 8, EXT_PARTS = 256;
constexpr int SS_WORDS = SS_PART + 8 * EXT_PARTS;

// min over the warp of a 64-bit key in two 32-bit reductions (high words, then the low words of the lanes that hold
// the winning high word)
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t k) {
  const uint32_t hi = (uint32_t)(k >> 32), lo = (uint32_t)k;
  const uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
  const uint32_t ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
  return ((uint64_t)mh << 32) | ml;
}
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t k) {
  const uint32_t hi = (uint32_t)(k >> 32), lo = (uint32_t)k;
  const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
  const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return ((uint64_t)mh << 32) | ml;
}
// Every thread of a 256-thread CTA contributes the position it just wrote (valid = false: none); the CTA folds its
// extreme keys into record blockIdx.x % EXT_PARTS (atomics spread over 256 records: no hot word).  The next build
// scales its 32-bit sort keys to the union of the records and derives the flat / planar flags from min == max
// (sort_prep) — the positions are in registers in those kernels, so the step needs no pass of its own over them (it
// used to: flat_detect, 24 bytes per particle and two more launches).
__device__ __forceinline__ void accumulate_extent(double x, double y, double z, bool valid, uint32_t light,
                                                  uint64_t* __restrict__ ss, uint64_t (*sm)[8]) {
  const uint64_t k[3] = {f64_key(x), f64_key(y), f64_key(z)};
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const uint64_t mn = warp_min_u64(valid ? k[d] : ~0ull), mx = warp_max_u64(valid ? k[d] : 0ull);
    if (lane == 0) sm[w][d] = mn, sm[w][3 + d] = mx;
  }
  const uint32_t wl = __reduce_max_sync(0xffffffffu, valid ? light : 0u);
  if (lane == 0) sm[w][6] = wl;
  __syncthreads();
  if (threadIdx.x < 7) {
    const int j = threadIdx.x;
    uint64_t v = sm[0][j];
    for (int q = 1; q < 8; ++q) v = j < 3 ? (sm[q][j] < v ? sm[q][j] : v) : (sm[q][j] > v ? sm[q][j] : v);
    unsigned long long* rec = reinterpret_cast<unsigned long long*>(ss + SS_PART + 8 * (blockIdx.x % EXT_PARTS));
    if (j < 3) atomicMin(rec + j, (unsigned long long)v);
    else atomicMax(rec + j, (unsigned long long)v);
  }
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
