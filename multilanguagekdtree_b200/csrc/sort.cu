// sort.cu — per-dimension sorted id lists: a hand-written stable LSD radix sort of the x, y and z coordinates, all
// three dimensions batched in one launch (blockIdx.y = dimension).  Starting from ids 0..n-1 and being stable, equal
// coordinates stay in ascending particle-id order: the canonical tie-break of the build (SURVEY.md §7 hard part 1).
//
// Two key widths share the kernels:
//  * 32-bit (4 passes x 8 bits, the path that normally runs): key32 = trunc((x - min) * (2^32 - 1) / (max - min)), a
//    MONOTONE (non-strict) map of the coordinate, so sorting by key32 orders everything except particles that share a
//    key32.  sort_fixup then orders every run of equal key32 by the full 64-bit key (runs of <= 32 entries, one thread
//    each; at N = 1M uniform over the extent a few hundred pairs collide) and checks every adjacent pair of longer runs;
//    a longer run that is out of order raises need64.
//  * 64-bit (8 passes x 8 bits over the order-preserving image of the f64): always correct; its 24 launches are
//    gated on need64 and return at once when the 32-bit result stands (clustered inputs whose extent / spacing
//    ratio exceeds 2^32, or non-finite coordinates, take this path).
// Either way the result is THE sorted order by (coordinate, id): bit-exact tree, 3.2x less sort traffic normally.
//
// This replaces, together with build.cu, the random-pivot quick-select of the reference
// (Parallel/RustVersion/src/quickstat.rs:9-34 called from array_kd_tree.rs:561-562): with the three lists
// sorted once per step, every node's median is the middle entry of its list segment and the bounding box
// is its two ends.
#include <algorithm>
#include <cstdlib>
#include <string>

#include "ctx.cuh"

namespace kdnb {

struct Pos3 {
  const double* p[3];
};

// (sort state layout: common.cuh)

__device__ __forceinline__ double key_to_f64(uint64_t k) {  // inverse of f64_key
  const uint64_t u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}
// the sort key of coordinate x in pass family K: the full order-preserving image, or its monotone 32-bit quantisation
template <typename K>
__device__ __forceinline__ K sort_key(double x, const uint64_t* __restrict__ ss, int d) {
  if (sizeof(K) == 8) return (K)f64_key(x);
  const double lo = __longlong_as_double((long long)ss[SS_LO + d]);
  const double scale = __longlong_as_double((long long)ss[SS_SCALE + d]);
  return (K)__double2uint_rz(__dmul_rn(__dsub_rn(x, lo), scale));  // saturating; monotone non-decreasing in x
}

// ---- pass kernel 1: per-tile digit histogram
template <bool FIRST, typename K>
__global__ void __launch_bounds__(SORT_THREADS) sort_upsweep(Pos3 pos, const K* __restrict__ keys_in,
                                                             uint32_t n, int shift, uint32_t ntiles,
                                                             uint32_t* __restrict__ hist,
                                                             const uint32_t* __restrict__ flat,
                                                             const uint64_t* __restrict__ ss, bool gated) {
  pdl_sync();
  __shared__ uint32_t h[256];
  const int d = blockIdx.y;
  if (flat[d]) return;  // all coordinates of this dimension are equal: its list is never consulted (build.cu)
  if (gated && !ss[SS_NEED64]) return;  // the 32-bit result stands
  const uint32_t tile = blockIdx.x;
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)tile * SORT_TILE;
#pragma unroll
  for (int k = 0; k < SORT_IPT; ++k) {
    uint64_t i = base + (uint64_t)k * SORT_THREADS + threadIdx.x;
    if (i < n) {
      K key = FIRST ? sort_key<K>(pos.p[d][i], ss, d) : keys_in[(uint64_t)d * n + i];
      atomicAdd(&h[(uint32_t)(key >> shift) & 255u], 1u);
    }
  }
  __syncthreads();
  hist[((uint64_t)d * 256 + threadIdx.x) * ntiles + tile] = h[threadIdx.x];
}

// ---- pass kernel 2: exclusive scan of every (dimension, digit) row over tiles; row totals to tot[]
__global__ void __launch_bounds__(256) sort_scan_rows(uint32_t* __restrict__ hist, uint32_t ntiles,
                                                      uint32_t* __restrict__ tot,
                                                      const uint32_t* __restrict__ flat,
                                                      const uint64_t* __restrict__ ss, bool gated) {
  pdl_sync();
  __shared__ uint32_t wsum[8];
  if (flat[blockIdx.y]) return;
  if (gated && !ss[SS_NEED64]) return;
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
template <bool FIRST, bool LAST, typename K>
__global__ void __launch_bounds__(SORT_THREADS, 8) sort_downsweep(Pos3 pos, const K* __restrict__ keys_in,
                                                               const uint32_t* __restrict__ vals_in,
                                                               K* __restrict__ keys_out,
                                                               uint32_t* __restrict__ vals_out, uint32_t n,
                                                               int shift, uint32_t ntiles,
                                                               const uint32_t* __restrict__ hist,
                                                               const uint32_t* __restrict__ tot,
                                                               const uint32_t* __restrict__ flat,
                                                               const uint64_t* __restrict__ ss, bool gated) {
  pdl_sync();
  __shared__ uint32_t wcnt[SORT_THREADS / 32][256];
  __shared__ uint32_t base[256];
  __shared__ uint32_t toff[256];
  __shared__ uint32_t wsum[8];
  __shared__ K skey[SORT_TILE];
  __shared__ uint32_t sval[SORT_TILE];
  const int d = blockIdx.y;
  if (flat[d]) return;
  if (gated && !ss[SS_NEED64]) return;
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

  K key[SORT_IPT];
  uint32_t val[SORT_IPT], rk[SORT_IPT];
  const uint64_t start = (uint64_t)tile * SORT_TILE + (uint64_t)w * (32 * SORT_IPT);
#pragma unroll
  for (int r = 0; r < SORT_IPT; ++r) {
    uint64_t i = start + r * 32 + lane;
    bool valid = i < n;
    key[r] = valid ? (FIRST ? sort_key<K>(pos.p[d][i], ss, d) : keys_in[(uint64_t)d * n + i]) : (K)~(K)0;
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
      const K kk = skey[p];
      const uint32_t dg = (uint32_t)((kk >> shift) & 255u);
      const uint64_t dst = (uint64_t)d * n + base[dg] + (p - toff[dg]);
      if (!LAST) keys_out[dst] = kk;
      vals_out[dst] = sval[p];
    }
  }
}

// sort_prep (one CTA) reduces — and resets — the extent records that the kernel which last wrote the positions
// accumulated (aos_to_soa, kick_drift: accumulate_extent, common.cuh):
//  * lo / scale of the 32-bit keys; non-finite extremes (inf or NaN coordinates) go straight to the 64-bit sort;
//  * flat[d] = 1 when every particle has the same coordinate d, i.e. min == max (planar inputs: the reference's ring has
//    z == 0 for ever).  Such a dimension has extent 0 in every node, so it is never the split dimension
//    (array_kd_tree.rs:551-556 keeps the lower dimension on ties) and its sorted list is never consulted: sorting and
//    partitioning it is skipped.  Dimension 0 is always kept — it is the tie winner when all extents are 0 and it
//    orders the leaves;
//  * flat[3] = 1 when, in addition, every z is +-0 and every mass is > 0 (the masses are seen at upload only): then every node's centre-of-mass z (sum m*z / sum m) is +-0 as well, dz == 0 in every test and
//    interaction, and the walk skips the z terms (walk2.cuh).
__global__ void __launch_bounds__(EXT_PARTS) sort_prep(uint64_t* ss, uint32_t* flat, uint32_t* dmask, int rank, int split_world,
                                                      int consume) {
  pdl_sync();
  __shared__ uint64_t sm[8][8];
  uint64_t v[7];
  if (consume) {  // (consume == 0: only the dimension shares of the split sort are re-derived, from the flags that stand)
  {
    uint64_t* rec = ss + SS_PART + 8 * threadIdx.x;  // one record per thread
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      v[j] = rec[j];
      rec[j] = j < 3 ? ~0ull : 0ull;
    }
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const uint64_t r = j < 3 ? warp_min_u64(v[j]) : warp_max_u64(v[j]);
    if (lane == 0) sm[w][j] = r;
  }
  __syncthreads();
  }
  if (threadIdx.x == 0) {
    uint64_t t[7];
    for (int j = 0; consume && j < 7; ++j) {
      t[j] = sm[0][j];
      for (int q = 1; q < 8; ++q) t[j] = j < 3 ? (sm[q][j] < t[j] ? sm[q][j] : t[j]) : (sm[q][j] > t[j] ? sm[q][j] : t[j]);
    }
    uint64_t need64 = 0ull;
    for (int d = 0; consume && d < 3; ++d) {
      const double lo = key_to_f64(t[d]), hi = key_to_f64(t[3 + d]);
      double scale = 0.0;  // all keys 0 when the extent is 0: one run of equal coordinates, already in id order
      if (!(fabs(lo) <= 1.7976931348623157e308) || !(fabs(hi) <= 1.7976931348623157e308)) need64 = 1ull;
      else if (hi > lo) scale = 4294967295.0 / (hi - lo);  // (hi - lo) may overflow to inf: scale 0, handled like extent 0 + fix-up
      ss[SS_LO + d] = (uint64_t)__double_as_longlong(lo);
      ss[SS_SCALE + d] = (uint64_t)__double_as_longlong(scale);
      flat[d] = (d > 0 && t[d] == t[3 + d]) ? 1u : 0u;
    }
    if (consume) {
      ss[SS_NEED64] = need64;
      if (t[6] < 2ull) ss[SS_LIGHT] = t[6];  // an upload's records (a kick leaves 2: the masses did not change)
      const uint64_t light = ss[SS_LIGHT];
      flat[3] = (t[2] == t[5] && t[2] == f64_key(0.0) && light == 0ull) ? 1u : 0u;
    }
    // split sort (split_world > 1): this rank sorts the non-flat dimensions whose index among them is congruent to its
    // rank modulo m = min(world, their number) and fetches every other one from the rank of its own block of m that
    // sorted it (sort_export / sort_fetch below)
    uint32_t nact = 0;
    for (int d = 0; d < 3; ++d) nact += flat[d] ? 0u : 1u;
    const uint32_t m = split_world > 1 ? min((uint32_t)split_world, nact) : 1u;
    uint32_t idx = 0;
    for (int d = 0; d < 3; ++d) {
      uint32_t skip = flat[d], src = 0xffffffffu;
      if (!flat[d]) {
        if (m > 1 && (idx % m) != ((uint32_t)rank % m)) {
          skip = 1u;
          src = (uint32_t)rank - (uint32_t)rank % m + idx % m;
          if (src >= (uint32_t)split_world) src = idx % m;
        }
        ++idx;
      }
      dmask[d] = skip;
      dmask[4 + d] = src;
    }
  }
}

// Split sort, donor side: the lists this rank sorted go to its export buffer (the list buffers themselves are permuted
// by the level partitions while peers would still be reading them); a system-scope fence and the last CTA then raise
// this rank's sort flag on every peer.
__global__ void __launch_bounds__(256) sort_export_kernel(P2PBuild pb, const uint32_t* __restrict__ lists, uint32_t n,
                                                          const uint32_t* __restrict__ dmask) {
  pdl_sync();
  const int me = pb.rank;
  uint32_t* out = pb.sexp[me];
  for (int d = 0; d < 3; ++d) {
    if (dmask[d]) continue;
    // (n is not a multiple of 4 in general and d * n not 16-byte aligned: vector body on the aligned part, scalar ends)
    const uint64_t base = (uint64_t)d * n;
    const uint64_t a0 = (4 - (base & 3)) & 3;  // elements until lists + base is 16-byte aligned
    const uint64_t nv = n > a0 ? (n - a0) / 4 : 0;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = tid; i < nv; i += nthr)
      reinterpret_cast<uint4*>(out + base + a0)[i] = reinterpret_cast<const uint4*>(lists + base + a0)[i];
    for (uint64_t i = tid; i < n; i += nthr)
      if (i < a0 || i >= a0 + 4 * nv) out[base + i] = lists[base + i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* st = pb.state[me];
    const uint32_t done = atomicAdd(&st[P2P_SDONE], 1u);
    if (done == gridDim.x - 1) {
      st[P2P_SDONE] = 0;
      __threadfence_system();
      const uint32_t epoch = st[P2P_SEPOCH];
      for (int p = 0; p < pb.world; ++p) *reinterpret_cast<volatile uint32_t*>(pb.state[p] + P2P_SFLAGS + me) = epoch + 1u;
    }
  }
}

// Split sort, receiving side: every dimension this rank did not sort is read from the export buffer of the peer that
// did, once that peer's flag says it is there.  The last CTA moves this rank's sort epoch on.  (No acknowledgement: a
// donor overwrites its export buffer in its NEXT sort, which starts after the step's all-rank exchange of the
// accelerations, and that exchange is behind this kernel on every rank.)
__global__ void __launch_bounds__(256) sort_fetch_kernel(P2PBuild pb, uint32_t* __restrict__ lists, uint32_t n,
                                                         const uint32_t* __restrict__ dmask) {
  pdl_sync();
  const int me = pb.rank;
  uint32_t* st = pb.state[me];
  const uint32_t target = st[P2P_SEPOCH] + 1u;
  for (int d = 0; d < 3; ++d) {
    const uint32_t src = dmask[4 + d];
    if (src == 0xffffffffu) continue;
    if (threadIdx.x == 0) {
      volatile uint32_t* flag = st + P2P_SFLAGS + src;
      const long long t0 = clock64();
      while (*flag < target) {
        if (clock64() - t0 > (30LL << 30)) {  // ~15 s: the donor died; record it instead of hanging for ever
          st[2] = 1u;
          break;
        }
      }
      __threadfence_system();
    }
    __syncthreads();
    const uint64_t base = (uint64_t)d * n;
    const uint32_t* in = pb.sexp[src] + base;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t a0 = (4 - (base & 3)) & 3;
    const uint64_t nv = n > a0 ? (n - a0) / 4 : 0;
    for (uint64_t i = tid; i < nv; i += nthr)
      reinterpret_cast<uint4*>(lists + base + a0)[i] = reinterpret_cast<const uint4*>(in + a0)[i];
    for (uint64_t i = tid; i < n; i += nthr)
      if (i < a0 || i >= a0 + 4 * nv) lists[base + i] = in[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t done = atomicAdd(&st[P2P_FDONE], 1u);
    if (done == gridDim.x - 1) {
      st[P2P_FDONE] = 0;
      st[P2P_SEPOCH] = target;
    }
  }
}

// Orders every run of equal key32 by (key64, id).  The 32-bit sort left each run in ascending id order, so a run is
// already right when its key64 are non-decreasing.  Run starts with <= FIX_CAP entries sort their run (stable insertion
// sort on key64, almost always a single compare of a pair); every other entry checks the pair (i-1, i) and, if that
// pair is out of order inside a run longer than FIX_CAP, requests the 64-bit sort.
constexpr int FIX_CAP = 32;
__global__ void __launch_bounds__(256) sort_fixup(Pos3 pos, const uint32_t* __restrict__ keys, uint32_t* __restrict__ lists,
                                                  uint32_t n, const uint32_t* __restrict__ flat,
                                                  uint64_t* __restrict__ ss, int kshift) {
  pdl_sync();
  const int d = blockIdx.y;
  if (flat[d] || ss[SS_NEED64]) return;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* key = keys + (uint64_t)d * n;
  uint32_t* lst = lists + (uint64_t)d * n;
  const double* x = pos.p[d];
  // (kshift > 0: the passes skipped the lowest digits — a run is a stretch of equal key >> kshift)
  const uint32_t k = key[i] >> kshift;
  const bool start = i == 0 || (key[i - 1] >> kshift) != k;
  if (start) {
    if (i + 1 >= n || (key[i + 1] >> kshift) != k) return;  // singleton
    uint32_t len = 2;
    while (len <= FIX_CAP && i + len < n && (key[i + len] >> kshift) == k) ++len;
    if (len > FIX_CAP) return;  // long run: verified pair by pair by its members
    if (len == 2) {
      const uint32_t a = lst[i], b = lst[i + 1];
      if (f64_key(x[a]) > f64_key(x[b])) lst[i] = b, lst[i + 1] = a;
      return;
    }
    uint32_t ids[FIX_CAP];
    uint64_t ks[FIX_CAP];
    bool moved = false;
    for (uint32_t r = 0; r < len; ++r) {
      const uint32_t id = lst[i + r];
      const uint64_t kk = f64_key(x[id]);
      int j = (int)r - 1;
      while (j >= 0 && ks[j] > kk) {
        ks[j + 1] = ks[j];
        ids[j + 1] = ids[j];
        --j;
        moved = true;
      }
      ks[j + 1] = kk;
      ids[j + 1] = id;
    }
    if (moved)
      for (uint32_t r = 0; r < len; ++r) lst[i + r] = ids[r];
  } else {
    const uint32_t a = lst[i - 1], b = lst[i];
    if (f64_key(x[a]) <= f64_key(x[b])) return;
    // out of order (or caught mid-update of a short run, which its start thread finishes): long run?
    uint32_t s = i, e = i + 1;
    while (s > 0 && i - s <= FIX_CAP && (key[s - 1] >> kshift) == k) --s;
    while (e < n && e - s <= FIX_CAP && (key[e] >> kshift) == k) ++e;
    if (e - s > FIX_CAP) ss[SS_NEED64] = 1ull;
  }
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

// hands need64 to the conditional node that holds the 64-bit passes (graph replay)
__global__ void sort_set_cond(cudaGraphConditionalHandle h, const uint64_t* __restrict__ ss) {
  pdl_sync();
  cudaGraphSetConditional(h, ss[SS_NEED64] ? 1u : 0u);
}

// Sorted lists end in c->list[0] whatever the number of passes: the buffers alternate so that the LAST pass writes
// buffer 0 (four passes: positions -> buf1 -> buf0 -> buf1 -> buf0; three: positions -> buf0 -> buf1 -> buf0).
// first_pass > 0 skips the lowest digits: the lists are then ordered by key >> (8 * first_pass) only and sort_fixup
// finishes the (longer) runs of equal truncated keys — see sort_lists().
template <typename K>
static void sort_passes(Ctx* c, Pos3 pos, bool gated, int first_pass = 0) {
  const uint32_t n = (uint32_t)c->n;
  const uint32_t nt = c->ntiles;
  const int passes = (int)sizeof(K);
  dim3 gt(nt, 3), gs(256, 3);
  K* kb[2] = {reinterpret_cast<K*>(c->keys[0]), reinterpret_cast<K*>(c->keys[1])};
  const uint64_t* ss = c->sort_state;
  for (int pass = first_pass; pass < passes; ++pass) {
    const int shift = 8 * pass;
    const int dst = ((passes - 1 - pass) & 1) ? 1 : 0, src = dst ^ 1;  // the last pass writes buffer 0
    const bool first = pass == first_pass;
    if (first) {
      KDNB_LAUNCH(c, (sort_upsweep<true, K>), gt, SORT_THREADS, 0, pos, nullptr, n, shift, nt, c->hist, c->dmask, ss, gated);
    } else {
      KDNB_LAUNCH(c, (sort_upsweep<false, K>), gt, SORT_THREADS, 0, pos, kb[src], n, shift, nt, c->hist, c->dmask, ss, gated);
    }
    KDNB_LAUNCH(c, sort_scan_rows, gs, 256, 0, c->hist, nt, c->digit_tot, c->dmask, ss, gated);
    if (first) {
      KDNB_LAUNCH(c, (sort_downsweep<true, false, K>), gt, SORT_THREADS, 0, pos, nullptr, nullptr, kb[dst], c->list[dst], n,
                  shift, nt, c->hist, c->digit_tot, c->dmask, ss, gated);
    } else if (pass == passes - 1 && sizeof(K) == 8) {  // (the 32-bit keys of the last pass are read by sort_fixup)
      KDNB_LAUNCH(c, (sort_downsweep<false, true, K>), gt, SORT_THREADS, 0, pos, kb[src], c->list[src], kb[dst],
                  c->list[dst], n, shift, nt, c->hist, c->digit_tot, c->dmask, ss, gated);
    } else {
      KDNB_LAUNCH(c, (sort_downsweep<false, false, K>), gt, SORT_THREADS, 0, pos, kb[src], c->list[src], kb[dst],
                  c->list[dst], n, shift, nt, c->hist, c->digit_tot, c->dmask, ss, gated);
    }
  }
}

// The 64-bit passes as the body of an IF node of the graph being captured on c->stream (condition: need64), so a
// replayed step pays for them only when they are needed.  Returns false when nothing was added (the caller then
// falls back on gated launches).
static bool sort64_conditional(Ctx* c, Pos3 pos) {
  if (!c->side_stream) return false;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaGraph_t g = nullptr;
  const cudaGraphNode_t* deps = nullptr;
  size_t nd = 0;
  if (cudaStreamGetCaptureInfo(c->stream, &st, nullptr, &g, &deps, &nd) != cudaSuccess ||
      st != cudaStreamCaptureStatusActive || !g) {
    cudaGetLastError();
    return false;
  }
  cudaGraphConditionalHandle h;
  if (cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  KDNB_LAUNCH(c, sort_set_cond, 1, 1, 0, h, (const uint64_t*)c->sort_state);
  if (cudaStreamGetCaptureInfo(c->stream, &st, nullptr, &g, &deps, &nd) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaGraphNodeParams p = {};
  p.type = cudaGraphNodeTypeConditional;
  p.conditional.handle = h;
  p.conditional.type = cudaGraphCondTypeIf;
  p.conditional.size = 1;
  cudaGraphNode_t node = nullptr;
  if (cudaGraphAddNode(&node, g, deps, nd, &p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaGraph_t body = p.conditional.phGraph_out[0];
  bool ok = cudaStreamBeginCaptureToGraph(c->side_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) ==
            cudaSuccess;
  if (ok) {
    cudaStream_t keep = c->stream;
    const uint64_t l0 = c->launches;
    c->stream = c->side_stream;
    sort_passes<uint64_t>(c, pos, false);
    c->stream = keep;
    c->launches = l0;  // not launched unless need64 is raised
    cudaGraph_t out = nullptr;
    ok = cudaStreamEndCapture(c->side_stream, &out) == cudaSuccess;
  }
  // the main capture continues after the conditional node (an empty body is harmless if its capture failed, but
  // then the gated launches must follow)
  if (cudaStreamUpdateCaptureDependencies(c->stream, &node, 1, cudaStreamSetCaptureDependencies) != cudaSuccess) ok = false;
  if (!ok) cudaGetLastError();
  return ok;
}

int sort_lists(Ctx* c) {
  const uint32_t n = (uint32_t)c->n;
  Pos3 pos = {{c->pos[0], c->pos[1], c->pos[2]}};
  // KDNB_SORT: "64" = skip the 32-bit path; "stubs" = always launch the 64-bit passes gated on need64
  static const char* knob = getenv("KDNB_SORT");
  static const bool only64 = knob && std::string(knob) == "64";
  static const bool stubs = knob && std::string(knob) == "stubs";
  // key scaling and flat / planar flags from the extents accumulated by whoever wrote the positions (upload or the
  // previous kick); a rebuild on unchanged positions keeps what the previous build derived
  // Split sort (multi-GPU, peer mode): every rank sorts a share of the dimensions and fetches the rest from a peer.
  // KDNB_SORT_SPLIT=0 disables, =1 forces it at every size.
  static const int split_mode = [] {
    const char* e = getenv("KDNB_SORT_SPLIT");
    return e ? (atoi(e) != 0 ? 1 : 0) : -1;
  }();
  const bool split = c->world > 1 && c->p2p_on && c->l0 > 0 && (split_mode == 1 || (split_mode < 0 && c->n >= (2ull << 20)));
  if (c->extent_fresh || split != c->split_sort)
    KDNB_LAUNCH(c, sort_prep, 1, EXT_PARTS, 0, c->sort_state, c->flat, c->dmask, c->rank_id, split ? c->world : 0,
                c->extent_fresh ? 1 : 0);
  c->split_sort = split;
  c->extent_fresh = false;
  if (only64) {
    sort_passes<uint64_t>(c, pos, false);
  } else {
    // Up to 4M particles THREE passes: the lowest digit is skipped, the lists come out ordered by the top 24 bits of
    // the scaled key and sort_fixup orders the runs of equal 24-bit keys (16.7M bins: a few per cent of the particles
    // share one at N = 1M; a run longer than 32 that is out of order still raises need64, i.e. only an input with
    // ~500 times the average density inside a 2^-24 slice of its extent falls back).  One radix pass fewer: -3 launches,
    // measured in profiles/r02_ab_sort_3pass.txt.  KDNB_SORT_PASSES=4 restores the four passes.
    static const bool four = [] {
      const char* e = getenv("KDNB_SORT_PASSES");
      return e && atoi(e) == 4;
    }();
    const int skip = (!four && c->n <= (4ull << 20)) ? 1 : 0;
    sort_passes<uint32_t>(c, pos, false, skip);
    KDNB_LAUNCH(c, sort_fixup, dim3((n + 255) / 256, 3), 256, 0, pos, reinterpret_cast<const uint32_t*>(c->keys[0]),
                c->list[0], n, c->dmask, c->sort_state, 8 * skip);
    // the 64-bit passes run only when need64 was raised: decided by the host between plain launches (one stream
    // synchronisation), by a conditional node inside a captured step, by the kernels themselves otherwise
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    KDNB_CUDA_TRY(c, cudaStreamIsCapturing(c->stream, &st));
    if (st == cudaStreamCaptureStatusActive) {
      if (stubs || c->use_pdl || !sort64_conditional(c, pos)) sort_passes<uint64_t>(c, pos, true);
    } else if (stubs || c->use_pdl) {
      sort_passes<uint64_t>(c, pos, true);
    } else {
      uint64_t need = 0;
      KDNB_CUDA_TRY(c, cudaMemcpyAsync(&need, c->sort_state + SS_NEED64, sizeof(need), cudaMemcpyDeviceToHost, c->stream));
      KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      if (need) sort_passes<uint64_t>(c, pos, false);
    }
  }
  if (split) {
    KDNB_LAUNCH(c, sort_export_kernel, 2 * c->num_sms, 256, 0, c->p2pb, c->list[0], n, c->dmask);
    KDNB_LAUNCH(c, sort_fetch_kernel, 2 * c->num_sms, 256, 0, c->p2pb, c->list[0], n, c->dmask);
  }
  if (c->l0 > 0) KDNB_LAUNCH(c, rank_from_lists, dim3((n + 255) / 256, 3), 256, 0, c->list[0], n, c->rk, c->flat);
  KDNB_CHECK_LAUNCH(c);
  return 0;
}

}  // namespace kdnb
