// sort.cu — per-dimension sorted id lists: a hand-written stable LSD radix sort (8 passes x 8 bits) of the
// order-preserving 64-bit keys of the x, y and z coordinates, all three dimensions batched in one launch
// (blockIdx.y = dimension).  Starting from ids 0..n-1 and being stable, equal coordinates stay in ascending
// particle-id order: the canonical tie-break of the build (SURVEY.md §7 hard part 1).
//
// This replaces, together with build.cu, the random-pivot quick-select of the reference
// (Parallel/RustVersion/src/quickstat.rs:9-34 called from array_kd_tree.rs:561-562): with the three lists
// sorted once per step, every node's median is the middle entry of its list segment and the bounding box
// is its two ends.
#include <algorithm>

#include "ctx.cuh"

namespace kdnb {

struct Pos3 {
  const double* p[3];
};

// ---- pass kernel 1: per-tile digit histogram
template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS) sort_upsweep(Pos3 pos, const uint64_t* __restrict__ keys_in,
                                                             uint32_t n, int shift, uint32_t ntiles,
                                                             uint32_t* __restrict__ hist,
                                                             const uint32_t* __restrict__ flat) {
  pdl_sync();
  __shared__ uint32_t h[256];
  const int d = blockIdx.y;
  if (flat[d]) return;  // all coordinates of this dimension are equal: its list is never consulted (build.cu)
  const uint32_t tile = blockIdx.x;
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)tile * SORT_TILE;
#pragma unroll
  for (int k = 0; k < SORT_IPT; ++k) {
    uint64_t i = base + (uint64_t)k * SORT_THREADS + threadIdx.x;
    if (i < n) {
      uint64_t key = FIRST ? f64_key(pos.p[d][i]) : keys_in[(uint64_t)d * n + i];
      atomicAdd(&h[(key >> shift) & 255u], 1u);
    }
  }
  __syncthreads();
  hist[((uint64_t)d * 256 + threadIdx.x) * ntiles + tile] = h[threadIdx.x];
}

// ---- pass kernel 2: exclusive scan of every (dimension, digit) row over tiles; row totals to tot[]
__global__ void __launch_bounds__(256) sort_scan_rows(uint32_t* __restrict__ hist, uint32_t ntiles,
                                                      uint32_t* __restrict__ tot,
                                                      const uint32_t* __restrict__ flat) {
  pdl_sync();
  __shared__ uint32_t wsum[8];
  if (flat[blockIdx.y]) return;
  uint32_t* row = hist + ((uint64_t)blockIdx.y * 256 + blockIdx.x) * ntiles;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < ntiles; base += 256) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < ntiles ? row[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    uint32_t wp = 0, total = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint32_t t = wsum[k];
      if (k < w) wp += t;
      total += t;
    }
    if (i < ntiles) row[i] = carry + wp + x - v;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) tot[blockIdx.y * 256 + blockIdx.x] = carry;
}

// ---- pass kernel 3: stable rank inside the tile (warp match), scatter
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(SORT_THREADS) sort_downsweep(Pos3 pos, const uint64_t* __restrict__ keys_in,
                                                               const uint32_t* __restrict__ vals_in,
                                                               uint64_t* __restrict__ keys_out,
                                                               uint32_t* __restrict__ vals_out, uint32_t n,
                                                               int shift, uint32_t ntiles,
                                                               const uint32_t* __restrict__ hist,
                                                               const uint32_t* __restrict__ tot,
                                                               const uint32_t* __restrict__ flat) {
  pdl_sync();
  __shared__ uint32_t wcnt[SORT_THREADS / 32][256];
  __shared__ uint32_t base[256];
  __shared__ uint32_t toff[256];
  __shared__ uint32_t wsum[8];
  __shared__ uint64_t skey[SORT_TILE];
  __shared__ uint32_t sval[SORT_TILE];
  const int d = blockIdx.y;
  if (flat[d]) return;
  const uint32_t tile = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;

#pragma unroll
  for (int k = 0; k < SORT_THREADS / 32; ++k) wcnt[k][threadIdx.x] = 0;
  {  // global base of every digit: exclusive scan of the 256 digit totals + this tile's row prefix
    uint32_t v = tot[d * 256 + threadIdx.x];
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    uint32_t wp = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < w) wp += wsum[k];
    base[threadIdx.x] = wp + x - v + hist[((uint64_t)d * 256 + threadIdx.x) * ntiles + tile];
  }
  __syncthreads();

  uint64_t key[SORT_IPT];
  uint32_t val[SORT_IPT], rk[SORT_IPT];
  const uint64_t start = (uint64_t)tile * SORT_TILE + (uint64_t)w * (32 * SORT_IPT);
#pragma unroll
  for (int r = 0; r < SORT_IPT; ++r) {
    uint64_t i = start + r * 32 + lane;
    bool valid = i < n;
    key[r] = valid ? (FIRST ? f64_key(pos.p[d][i]) : keys_in[(uint64_t)d * n + i]) : ~0ull;
    val[r] = valid ? (FIRST ? (uint32_t)i : vals_in[(uint64_t)d * n + i]) : 0u;
    uint32_t dg = valid ? (uint32_t)((key[r] >> shift) & 255u) : 256u;
    uint32_t peers = __match_any_sync(0xffffffffu, dg);
    int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (valid && lane == leader) {
      old = wcnt[w][dg];
      wcnt[w][dg] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rk[r] = old + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  {  // exclusive scan over warps, per digit; then exclusive scan over digits = start of each digit's run in the tile
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < SORT_THREADS / 32; ++k) {
      uint32_t t = wcnt[k][threadIdx.x];
      wcnt[k][threadIdx.x] = run;
      run += t;
    }
    uint32_t x = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    uint32_t wp = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < w) wp += wsum[k];
    toff[threadIdx.x] = wp + x - run;
  }
  __syncthreads();
  // stage the tile in digit order in shared memory, then write it out in runs: consecutive threads write consecutive
  // addresses inside each digit's run (the direct scatter wrote 12-byte granules all over the output)
#pragma unroll
  for (int r = 0; r < SORT_IPT; ++r) {
    uint64_t i = start + r * 32 + lane;
    if (i < n) {
      uint32_t dg = (uint32_t)((key[r] >> shift) & 255u);
      uint32_t p = toff[dg] + wcnt[w][dg] + rk[r];
      skey[p] = key[r];
      sval[p] = val[r];
    }
  }
  __syncthreads();
  const uint64_t tile_base = (uint64_t)tile * SORT_TILE;
  const uint32_t nvalid = (uint32_t)min((uint64_t)SORT_TILE, (uint64_t)n - tile_base);
#pragma unroll
  for (int k = 0; k < SORT_IPT; ++k) {
    const uint32_t p = k * SORT_THREADS + threadIdx.x;
    if (p < nvalid) {
      const uint64_t kk = skey[p];
      const uint32_t dg = (uint32_t)((kk >> shift) & 255u);
      const uint64_t dst = (uint64_t)d * n + base[dg] + (p - toff[dg]);
      if (!LAST) keys_out[dst] = kk;
      vals_out[dst] = sval[p];
    }
  }
}

// flat[d] = 1 when every particle has the same coordinate d (planar inputs: the reference's ring has z == 0 for
// ever).  Such a dimension has extent 0 in every node, so it is never the split dimension (array_kd_tree.rs:551-556
// keeps the lower dimension on ties) and its sorted list is never consulted: sorting and partitioning it is skipped.
// Dimension 0 is always kept — it is the tie winner when all extents are 0 and it orders the leaves.
__global__ void flat_init(uint32_t* flat) {
  pdl_sync();
  if (threadIdx.x < 4) flat[threadIdx.x] = threadIdx.x > 0 ? 1u : 0u;
}
// flat[3] = 1 when, in addition, every z is +-0 and every mass is > 0: then every node's centre-of-mass z
// (sum m*z / sum m) is +-0 as well, dz == 0 in every test and interaction, and the walk skips the z terms (walk2.cuh).
__global__ void __launch_bounds__(256) flat_detect(Pos3 pos, const double* __restrict__ mass, uint32_t n,
                                                   uint32_t* __restrict__ flat) {
  pdl_sync();
  const int d = blockIdx.y + 1;
  const uint64_t k0 = f64_key(pos.p[d][0]);
  bool differs = false, heavy = true;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    differs |= f64_key(pos.p[d][i]) != k0;
    if (d == 2) heavy &= mass[i] > 0.0;
  }
  if (differs) flat[d] = 0u;
  if (d == 2 && (differs || !heavy || k0 != f64_key(0.0))) flat[3] = 0u;
}

// rk[d][id] = rank of particle id in the sorted list of dimension d (read by the global partition levels, build.cu)
__global__ void __launch_bounds__(256) rank_from_lists(const uint32_t* __restrict__ lists, uint32_t n,
                                                       uint32_t* __restrict__ rk, const uint32_t* __restrict__ flat) {
  pdl_sync();
  const int d = blockIdx.y;
  if (flat[d]) return;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rk[(uint64_t)d * n + lists[(uint64_t)d * n + i]] = i;
}

// Sorted lists end in c->list[0] (8 passes: positions -> buf1 -> buf0 -> ... -> buf0).
int sort_lists(Ctx* c) {
  const uint32_t n = (uint32_t)c->n;
  const uint32_t nt = c->ntiles;
  Pos3 pos = {{c->pos[0], c->pos[1], c->pos[2]}};
  dim3 gt(nt, 3), gs(256, 3);
  KDNB_LAUNCH(c, flat_init, 1, 32, 0, c->flat);
  KDNB_LAUNCH(c, flat_detect, dim3(std::min<uint32_t>((n + 255) / 256, 1184u), 2), 256, 0, pos, c->mass, n, c->flat);
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 8 * pass;
    const int src = (pass & 1) ? 1 : 0, dst = src ^ 1;  // pass 0 reads positions, writes buf1
    if (pass == 0) {
      KDNB_LAUNCH(c, sort_upsweep<true>, gt, SORT_THREADS, 0, pos, nullptr, n, shift, nt, c->hist, c->flat);
    } else {
      KDNB_LAUNCH(c, sort_upsweep<false>, gt, SORT_THREADS, 0, pos, c->keys[src], n, shift, nt, c->hist, c->flat);
    }
    KDNB_LAUNCH(c, sort_scan_rows, gs, 256, 0, c->hist, nt, c->digit_tot, c->flat);
    if (pass == 0) {
      KDNB_LAUNCH(c, (sort_downsweep<true, false>), gt, SORT_THREADS, 0, pos, nullptr, nullptr, c->keys[1],
                  c->list[1], n, shift, nt, c->hist, c->digit_tot, c->flat);
    } else if (pass == 7) {
      KDNB_LAUNCH(c, (sort_downsweep<false, true>), gt, SORT_THREADS, 0, pos, c->keys[src], c->list[src],
                  c->keys[dst], c->list[dst], n, shift, nt, c->hist, c->digit_tot, c->flat);
    } else {
      KDNB_LAUNCH(c, (sort_downsweep<false, false>), gt, SORT_THREADS, 0, pos, c->keys[src], c->list[src],
                  c->keys[dst], c->list[dst], n, shift, nt, c->hist, c->digit_tot, c->flat);
    }
  }
  if (c->l0 > 0) KDNB_LAUNCH(c, rank_from_lists, dim3((n + 255) / 256, 3), 256, 0, c->list[0], n, c->rk, c->flat);
  KDNB_CHECK_LAUNCH(c);
  return 0;
}

}  // namespace kdnb
