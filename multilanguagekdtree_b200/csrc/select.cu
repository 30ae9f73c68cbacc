// select.cu — quickstat_index of the reference (Parallel/RustVersion/src/quickstat.rs:9-34, benchmarked stand-alone by
// src/bin/bench_quickstat.rs on 100M values) on the device.
//
// The reference permutes `indices` in place with a random-pivot quick-select until position `goal` holds the
// goal-th smallest element under `lt`; which permutation comes out depends on the pivots, and the post-condition its
// tests check (quickstat.rs:199-253) is
//      for i < goal: !(v[indices[goal]] < v[indices[i]])      for i > goal: !(v[indices[i]] < v[indices[goal]]).
// Here: a most-significant-byte-first RADIX SELECT over the order-preserving 64-bit image of the f64 values (eight
// histogram passes, each narrowing the candidate prefix by one byte, no data movement) finds the key of the goal-th
// smallest value exactly; one stable three-way partition (< pivot | == pivot | > pivot) then produces the permutation.
// Position `goal` falls inside the == block, so the post-condition holds, and among the permutations the reference
// may produce this is the canonical one: every block keeps the input order (what the tree build's tie-break needs).
// HBM traffic: 8 B/element for the keys once, 8 passes x 8 B, partition 2 x (8 + 4) B: ~100 B/element, ~2 ms at 100M.
#include <algorithm>

#include "ctx.cuh"

namespace kdnb {

constexpr int SEL_THREADS = 256;
constexpr int SEL_IPT = 8;
constexpr int SEL_TILE = SEL_THREADS * SEL_IPT;

// selection state (device, u64 words): [0] prefix value, [1] prefix mask, [2] remaining goal inside the prefix,
// [3] number of keys < pivot, [4] number of keys == pivot; histogram (256 x u32) behind it
constexpr int SL_PREFIX = 0, SL_MASK = 1, SL_GOAL = 2, SL_LESS = 3, SL_EQUAL = 4, SL_WORDS = 8;

__global__ void __launch_bounds__(SEL_THREADS) sel_keys(const double* __restrict__ vals, const uint64_t* __restrict__ idx64,
                                                        uint64_t count, uint64_t* __restrict__ keys,
                                                        uint32_t* __restrict__ idx32) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t id = idx64 ? idx64[i] : i;
    keys[i] = f64_key(vals[id]);
    idx32[i] = (uint32_t)id;
  }
}

__global__ void sel_init(uint64_t* st, uint32_t* hist, uint64_t goal) {
  if (threadIdx.x == 0) {
    st[SL_PREFIX] = 0, st[SL_MASK] = 0, st[SL_GOAL] = goal, st[SL_LESS] = 0, st[SL_EQUAL] = 0;
  }
  hist[threadIdx.x] = 0;
}

// histogram of byte `shift/8` over the keys that match the prefix found so far
__global__ void __launch_bounds__(SEL_THREADS) sel_hist(const uint64_t* __restrict__ keys, uint64_t count, int shift,
                                                        const uint64_t* __restrict__ st, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t prefix = st[SL_PREFIX], mask = st[SL_MASK];
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    if ((k & mask) == prefix) atomicAdd(&h[(uint32_t)(k >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

// the bucket that holds the goal: extend the prefix by one byte
__global__ void sel_pick(uint64_t* st, uint32_t* hist, int shift) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = hist[threadIdx.x];
  hist[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t goal = st[SL_GOAL], before = 0;
    int b = 0;
    for (; b < 255; ++b) {
      if (before + h[b] > goal) break;
      before += h[b];
    }
    st[SL_GOAL] = goal - before;
    st[SL_LESS] += before;  // keys with a smaller byte under the same prefix are smaller than the pivot
    st[SL_PREFIX] |= (uint64_t)b << shift;
    st[SL_MASK] |= 255ull << shift;
    if (shift == 0) st[SL_EQUAL] = h[b];
  }
}

// stable three-way partition: per-tile counts, one-CTA scan of the tile counts, scatter
__global__ void __launch_bounds__(SEL_THREADS) sel_count(const uint64_t* __restrict__ keys, uint64_t count,
                                                         const uint64_t* __restrict__ st, uint2* __restrict__ tile_cnt) {
  __shared__ uint32_t wl[SEL_THREADS / 32], we[SEL_THREADS / 32];
  const uint64_t pivot = st[SL_PREFIX];
  const uint64_t base = (uint64_t)blockIdx.x * SEL_TILE;
  uint32_t nl = 0, ne = 0;
#pragma unroll
  for (int k = 0; k < SEL_IPT; ++k) {
    const uint64_t i = base + (uint64_t)k * SEL_THREADS + threadIdx.x;
    if (i < count) {
      const uint64_t key = keys[i];
      nl += key < pivot;
      ne += key == pivot;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nl += __shfl_xor_sync(0xffffffffu, nl, o);
    ne += __shfl_xor_sync(0xffffffffu, ne, o);
  }
  if ((threadIdx.x & 31) == 0) wl[threadIdx.x >> 5] = nl, we[threadIdx.x >> 5] = ne;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t a = 0, b = 0;
    for (int k = 0; k < SEL_THREADS / 32; ++k) a += wl[k], b += we[k];
    tile_cnt[blockIdx.x] = make_uint2(a, b);
  }
}

__global__ void __launch_bounds__(1024) sel_scan(uint2* __restrict__ tile_cnt, uint32_t ntiles) {
  __shared__ uint32_t sa[32], sb[32];
  __shared__ uint32_t ca, cb;
  if (threadIdx.x == 0) ca = cb = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint32_t base = 0; base < ntiles; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint2 v = i < ntiles ? tile_cnt[i] : make_uint2(0u, 0u);
    uint32_t a = v.x, b = v.y;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t ya = __shfl_up_sync(0xffffffffu, a, o), yb = __shfl_up_sync(0xffffffffu, b, o);
      if (lane >= o) a += ya, b += yb;
    }
    if (lane == 31) sa[w] = a, sb[w] = b;
    __syncthreads();
    uint32_t pa = ca, pb = cb;
    for (int k = 0; k < w; ++k) pa += sa[k], pb += sb[k];
    if (i < ntiles) tile_cnt[i] = make_uint2(pa + a - v.x, pb + b - v.y);  // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) ca = pa + a, cb = pb + b;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SEL_THREADS) sel_scatter(const uint64_t* __restrict__ keys,
                                                           const uint32_t* __restrict__ idx_in, uint64_t count,
                                                           const uint64_t* __restrict__ st,
                                                           const uint2* __restrict__ tile_off,
                                                           uint64_t* __restrict__ idx_out) {
  __shared__ uint32_t wl[SEL_THREADS / 32], we[SEL_THREADS / 32];
  const uint64_t pivot = st[SL_PREFIX], nless = st[SL_LESS], nequal = st[SL_EQUAL];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  // warp w owns SEL_IPT consecutive groups of 32 (input order = tile order = output order inside every block)
  const uint64_t wbase = (uint64_t)blockIdx.x * SEL_TILE + (uint64_t)w * (32 * SEL_IPT);
  uint32_t bl[SEL_IPT], be[SEL_IPT], id[SEL_IPT];
  uint32_t nl = 0, ne = 0;
#pragma unroll
  for (int k = 0; k < SEL_IPT; ++k) {
    const uint64_t i = wbase + k * 32 + lane;
    uint64_t key = ~0ull;
    bool valid = i < count;
    if (valid) key = keys[i], id[k] = idx_in[i];
    bl[k] = __ballot_sync(0xffffffffu, valid && key < pivot);
    be[k] = __ballot_sync(0xffffffffu, valid && key == pivot);
    nl += __popc(bl[k]);
    ne += __popc(be[k]);
  }
  if (lane == 0) wl[w] = nl, we[w] = ne;
  __syncthreads();
  const uint2 off = tile_off[blockIdx.x];
  uint64_t pl = off.x, pe = off.y;
  for (int k = 0; k < w; ++k) pl += wl[k], pe += we[k];
#pragma unroll
  for (int k = 0; k < SEL_IPT; ++k) {
    const uint64_t i = wbase + k * 32 + lane;
    if (i < count) {
      const bool less = (bl[k] >> lane) & 1u, equal = (be[k] >> lane) & 1u;
      const uint64_t l = pl + __popc(bl[k] & lt), e = pe + __popc(be[k] & lt);
      const uint64_t dst = less ? l : equal ? nless + e : nless + nequal + (i - l - e);
      idx_out[dst] = id[k];
    }
    pl += __popc(bl[k]);
    pe += __popc(be[k]);
  }
}

// vals / idx64 / keys / idx32 / out64 / tile_cnt are scratch device buffers of the call; st + hist live in `state`
int select_run(Ctx* c, const double* d_vals, const uint64_t* d_idx64, uint64_t count, uint64_t goal, uint64_t* d_keys,
               uint32_t* d_idx32, uint2* d_tiles, uint64_t* d_state, uint64_t* d_out64) {
  uint32_t* hist = reinterpret_cast<uint32_t*>(d_state + SL_WORDS);
  const uint32_t ntiles = (uint32_t)((count + SEL_TILE - 1) / SEL_TILE);
  const uint32_t grid = std::min<uint32_t>(ntiles, (uint32_t)c->num_sms * 16u);
  KDNB_LAUNCH(c, sel_keys, grid, SEL_THREADS, 0, d_vals, d_idx64, count, d_keys, d_idx32);
  KDNB_LAUNCH(c, sel_init, 1, 256, 0, d_state, hist, goal);
  for (int pass = 7; pass >= 0; --pass) {
    KDNB_LAUNCH(c, sel_hist, grid, SEL_THREADS, 0, (const uint64_t*)d_keys, count, 8 * pass, (const uint64_t*)d_state, hist);
    KDNB_LAUNCH(c, sel_pick, 1, 256, 0, d_state, hist, 8 * pass);
  }
  KDNB_LAUNCH(c, sel_count, ntiles, SEL_THREADS, 0, (const uint64_t*)d_keys, count, (const uint64_t*)d_state, d_tiles);
  KDNB_LAUNCH(c, sel_scan, 1, 1024, 0, d_tiles, ntiles);
  KDNB_LAUNCH(c, sel_scatter, ntiles, SEL_THREADS, 0, (const uint64_t*)d_keys, (const uint32_t*)d_idx32, count,
              (const uint64_t*)d_state, (const uint2*)d_tiles, d_out64);
  KDNB_CHECK_LAUNCH(c);
  return 0;
}

}  // namespace kdnb
