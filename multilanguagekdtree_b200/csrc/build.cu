// build.cu — level-synchronous kD-tree build reproducing build_tree_par4 / build_tree
// (Parallel/RustVersion/src/array_kd_tree.rs:515-583 and :63-130) node for node.
//
// Input: the three per-dimension sorted id lists from sort.cu.  Invariant kept at every level: inside each
// node's slot range [a, a+len) every list holds exactly that node's particles, sorted by its dimension
// (ties by ascending id).  Then for a node
//   bbox[d]   = coordinates of the first / last entry of list d        (array_kd_tree.rs:532-547, min/max part)
//   split_dim = widest extent, strict '>' so ties keep the lower dim    (:551-556)
//   size      = extent along split_dim                                  (:557)
//   mid       = a + len/2, split_val = coordinate of list[split_dim][mid] (:560-563)
// and the children are obtained by a STABLE partition of the other two lists by "is in the left half of
// list[split_dim]" — no selection passes, no reductions.
//
//  * levels whose segments are larger than BOT_CAP run as global kernels (stats, flags, count, scan, scatter);
//  * once segments fit (<= BOT_CAP slots) ONE kernel finishes all remaining levels in shared memory, emits the
//    leaves, the tree-ordered particle copies and sums m / sum(m*p) bottom-up inside the segment;
//  * a last single-CTA kernel carries m / sum(m*p) up the few global levels.
// m and cm are summed in the canonical order (leaf: ascending id, internal: left + right), see DESIGN.md.
#include <cstdlib>

#include "ctx.cuh"

namespace kdnb {

struct Lists {
  uint32_t* l[3];
};
struct Pos3c {
  const double* p[3];
};

// ------------------------------------------------------------------------------------------ global levels

// lvl_ctl (device): [0..63] per-level CTA tickets, [64] build epoch, [65] look-back timeout flag
constexpr int LC_EPOCH = 64, LC_ERR = 65, LC_WORDS = 72;

__global__ void build_root(uint32_t* tstart, uint32_t* tlen, uint32_t* tnode, uint32_t n, uint32_t* lvl_ctl) {
  pdl_sync();
  if (threadIdx.x == 0) {
    tstart[0] = 0;
    tlen[0] = n;
    tnode[0] = 0;
    lvl_ctl[LC_EPOCH] += 1u;
  }
  if (threadIdx.x < 64) lvl_ctl[threadIdx.x] = 0u;
}

// Statistics of segment s of `level` (bbox from the list ends, split dimension, median).  Every CTA working on the
// segment evaluates this (a handful of gathers); the one flagged `writer` also emits the node record and the table
// entries of the two children.  Returns sd, mid and the initial rank of the median element.
struct SegStats {
  uint32_t a, len, sd, mid, rmid;
};
__device__ __forceinline__ SegStats seg_stats(Pos3c pos, Lists L, int level, uint32_t s, uint32_t mp, int layout,
                                              uint32_t* __restrict__ tstart, uint32_t* __restrict__ tlen,
                                              uint32_t* __restrict__ tnode, uint32_t* __restrict__ tmid,
                                              uint8_t* __restrict__ tsd, WNode* __restrict__ nodes,
                                              const uint32_t* __restrict__ flat, const uint32_t* __restrict__ rk,
                                              uint32_t n, uint32_t* __restrict__ tmr, bool writer) {
  const uint32_t nseg = 1u << level;
  const uint32_t off = nseg - 1;
  const uint32_t a = tstart[off + s], len = tlen[off + s], node = tnode[off + s];
  double mn[3], mx[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    mn[d] = mx[d] = 0.0;  // a flat dimension has extent 0 everywhere (its list is not maintained)
    if (!flat[d]) {
      mn[d] = pos.p[d][L.l[d][a]];
      mx[d] = pos.p[d][L.l[d][a + len - 1]];
    }
  }
  int sd = 0;
  double ext = mx[0] - mn[0];
#pragma unroll
  for (int d = 1; d < 3; ++d) {
    double e = mx[d] - mn[d];
    if (e > ext) {
      ext = e;
      sd = d;
    }
  }
  const uint32_t half = len / 2, mid = a + half;
  const uint32_t mid_id = L.l[sd][mid];
  SegStats r;
  r.a = a;
  r.len = len;
  r.sd = (uint32_t)sd;
  r.mid = mid;
  r.rmid = rk[(uint64_t)sd * n + mid_id];
  if (writer) {
    const double split_val = pos.p[sd][mid_id];
    const uint32_t nleft = (uint32_t)subtree_nodes(half, mp, layout);
    WNode* nd = &nodes[node];
    nd->size2 = __dmul_rn(ext, ext);
    nd->size = ext;
    nd->split_val = split_val;
    nd->a = node + 1 + nleft;
    nd->b = WN_INTERNAL | (uint32_t)sd;
    tsd[off + s] = (uint8_t)sd;
    tmid[off + s] = mid;
    tmr[off + s] = r.rmid;
    const uint32_t coff = 2 * nseg - 1;
    tstart[coff + 2 * s] = a;
    tlen[coff + 2 * s] = half;
    tnode[coff + 2 * s] = node + 1;
    tstart[coff + 2 * s + 1] = mid;
    tlen[coff + 2 * s + 1] = len - half;
    tnode[coff + 2 * s + 1] = node + 1 + nleft;
  }
  return r;
}

// "goes left" is decided without any per-level flag pass: rk[d][id] is the rank of particle id in the INITIAL sorted
// list of dimension d (sort.cu).  Stable partitions keep every segment of list d ordered by rk[d], so the left half of
// a node split along sd is exactly { id : rk[sd][id] < rk[sd][id of the element at mid] } (tmr[] holds that rank).
__global__ void __launch_bounds__(LVL_THREADS)
level_count(Pos3c pos, Lists L, int level, uint32_t cps, uint32_t mp, int layout, uint32_t* __restrict__ tstart,
            uint32_t* __restrict__ tlen, uint32_t* __restrict__ tnode, uint32_t* __restrict__ tmid,
            uint8_t* __restrict__ tsd, WNode* __restrict__ nodes, const uint32_t* __restrict__ rk, uint32_t n,
            uint32_t* __restrict__ tmr, uint32_t* __restrict__ cnt, const uint32_t* __restrict__ flat) {
  pdl_sync();
  __shared__ uint32_t wsum[LVL_THREADS / 32];
  __shared__ SegStats st;
  const uint32_t seg = blockIdx.x / cps, chunk = blockIdx.x % cps, e = blockIdx.y;
  const uint32_t nseg = 1u << level;
  if (flat[e]) return;  // (dimension 0 is never flat: its chunk-0 CTA is the writer)
  if (threadIdx.x == 0)
    st = seg_stats(pos, L, level, seg, mp, layout, tstart, tlen, tnode, tmid, tsd, nodes, flat, rk, n, tmr,
                   chunk == 0 && e == 0);
  __syncthreads();
  if (st.sd == e) return;  // the split-dimension list is already partitioned
  const uint32_t a = st.a, len = st.len, rmid = st.rmid;
  const uint32_t* lst = L.l[e];
  const uint32_t* rks = rk + (uint64_t)st.sd * n;
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < LVL_CHUNK / LVL_THREADS; ++k) {
    uint32_t o = chunk * LVL_CHUNK + k * LVL_THREADS + threadIdx.x;
    if (o < len) c += (rks[lst[a + o]] < rmid);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int k = 0; k < LVL_THREADS / 32; ++k) t += wsum[k];
    cnt[((uint64_t)e * nseg + seg) * cps + chunk] = t;
  }
}

// stable partition of list e inside each segment (copy for the split-dimension list)
__global__ void __launch_bounds__(LVL_THREADS, 8) level_scatter(Lists Lin, Lists Lout, int level, uint32_t cps,
                                                             const uint32_t* __restrict__ tstart,
                                                             const uint32_t* __restrict__ tlen,
                                                             const uint32_t* __restrict__ tmid,
                                                             const uint8_t* __restrict__ tsd,
                                                             const uint32_t* __restrict__ rk, uint32_t n,
                                                             const uint32_t* __restrict__ tmr,
                                                             const uint32_t* __restrict__ cnt,
                                                             const uint32_t* __restrict__ flat) {
  pdl_sync();
  constexpr int IPT = LVL_CHUNK / LVL_THREADS;  // 8
  __shared__ uint32_t wtot[LVL_THREADS / 32];
  const uint32_t seg = blockIdx.x / cps, chunk = blockIdx.x % cps, e = blockIdx.y;
  if (flat[e]) return;
  const uint32_t nseg = 1u << level, off = nseg - 1;
  const uint32_t a = tstart[off + seg], len = tlen[off + seg], mid = tmid[off + seg];
  const uint32_t* lin = Lin.l[e];
  uint32_t* lout = Lout.l[e];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t wbase_off = chunk * LVL_CHUNK + w * (32 * IPT);

  if (tsd[off + seg] == e) {
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      uint32_t o = wbase_off + k * 32 + lane;
      if (o < len) lout[a + o] = lin[a + o];
    }
    return;
  }
  const uint32_t* rks = rk + (uint64_t)tsd[off + seg] * n;
  const uint32_t rmid = tmr[off + seg];
  uint32_t id[IPT], bl[IPT];
  uint32_t wl = 0;
#pragma unroll
  for (int k = 0; k < IPT; ++k) {
    uint32_t o = wbase_off + k * 32 + lane;
    bool valid = o < len;
    id[k] = valid ? lin[a + o] : 0u;
    bool isleft = valid && (rks[id[k]] < rmid);
    bl[k] = __ballot_sync(0xffffffffu, isleft);
    wl += __popc(bl[k]);
  }
  if (lane == 0) wtot[w] = wl;
  __syncthreads();
  uint32_t wbase = 0;
#pragma unroll
  for (int k = 0; k < LVL_THREADS / 32; ++k)
    if (k < w) wbase += wtot[k];
  // lefts in the earlier chunks of this segment: block-wide sum of their counts (replaces a separate scan kernel)
  __shared__ uint32_t lbsum[LVL_THREADS / 32];
  uint32_t lb = 0;
  {
    const uint32_t* row = cnt + ((uint64_t)e * nseg + seg) * cps;
    for (uint32_t k = threadIdx.x; k < chunk; k += LVL_THREADS) lb += row[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lb += __shfl_xor_sync(0xffffffffu, lb, o);
    if (lane == 0) lbsum[w] = lb;
    __syncthreads();
    lb = 0;
#pragma unroll
    for (int k = 0; k < LVL_THREADS / 32; ++k) lb += lbsum[k];
  }
  const uint32_t leftbase = lb;
  uint32_t pre = wbase;
#pragma unroll
  for (int k = 0; k < IPT; ++k) {
    uint32_t o = wbase_off + k * 32 + lane;
    if (o < len) {
      uint32_t lrank = leftbase + pre + __popc(bl[k] & lt);  // lefts before me in the segment
      bool isleft = (bl[k] >> lane) & 1u;
      uint32_t dst = isleft ? a + lrank : mid + (o - lrank);
      lout[dst] = id[k];
    }
    pre += __popc(bl[k]);
  }
}

// ---- one kernel per global level: stable partition of every list inside every segment, single pass.
// A CTA owns one chunk (LVL_CHUNK entries) of one segment of one list.  It takes a ticket (CTAs are numbered in the
// order they START, so a CTA only ever waits for CTAs that are already running), evaluates the segment statistics,
// flags its entries (left = initial rank along the split dimension below the median's), and obtains the number of
// lefts in the earlier chunks of its segment by decoupled look-back over per-chunk status words
//     epoch (30 bits) | state (2 bits: 1 = chunk count, 2 = inclusive prefix) | value (32 bits)
// written and polled as single 64-bit words.  The epoch (build counter * 64 + level + 1) makes stale words of earlier
// launches unreadable, so the array is never cleared.  This replaces the count + scatter kernel pair (which read the
// lists and gathered the ranks twice and paid two launch latencies per level).
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(LVL_THREADS)
level_partition(Pos3c pos, Lists Lin, Lists Lout, int level, uint32_t cps, uint32_t mp, int layout,
                uint32_t* __restrict__ tstart, uint32_t* __restrict__ tlen, uint32_t* __restrict__ tnode,
                uint32_t* __restrict__ tmid, uint8_t* __restrict__ tsd, WNode* __restrict__ nodes,
                const uint32_t* __restrict__ rk, uint32_t n, uint32_t* __restrict__ tmr,
                unsigned long long* __restrict__ status, uint32_t* __restrict__ lvl_ctl,
                const uint32_t* __restrict__ flat) {
  pdl_sync();
  constexpr int IPT = LVL_CHUNK / LVL_THREADS;  // 8
  __shared__ uint32_t wtot[LVL_THREADS / 32];
  __shared__ SegStats st;
  __shared__ uint32_t s_ticket, s_leftbase;
  if (threadIdx.x == 0) s_ticket = atomicAdd(&lvl_ctl[level], 1u);
  __syncthreads();
  const uint32_t nseg = 1u << level, off = nseg - 1, gx = nseg * cps;
  const uint32_t e = s_ticket / gx, bx = s_ticket % gx;
  const uint32_t seg = bx / cps, chunk = bx % cps;
  if (flat[e]) return;  // (dimension 0 is never flat: its chunk-0 CTA is the writer)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  if (threadIdx.x == 0)
    st = seg_stats(pos, Lin, level, seg, mp, layout, tstart, tlen, tnode, tmid, tsd, nodes, flat, rk, n, tmr,
                   chunk == 0 && e == 0);
  // the entries of this chunk are loaded while thread 0 walks the dependent loads of the statistics
  const uint32_t a = tstart[off + seg], len = tlen[off + seg];
  const uint32_t* lin = Lin.l[e];
  uint32_t* lout = Lout.l[e];
  const uint32_t wbase_off = chunk * LVL_CHUNK + w * (32 * IPT);
  uint32_t id[IPT];
#pragma unroll
  for (int k = 0; k < IPT; ++k) {
    const uint32_t o = wbase_off + k * 32 + lane;
    id[k] = o < len ? lin[a + o] : 0u;
  }
  __syncthreads();
  if (st.sd == e) {  // the split-dimension list is already partitioned: copy
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const uint32_t o = wbase_off + k * 32 + lane;
      if (o < len) lout[a + o] = id[k];
    }
    return;
  }
  const uint32_t* rks = rk + (uint64_t)st.sd * n;
  const uint32_t rmid = st.rmid, mid = st.mid;
  uint32_t bl[IPT];
  uint32_t wl = 0;
#pragma unroll
  for (int k = 0; k < IPT; ++k) {
    const uint32_t o = wbase_off + k * 32 + lane;
    const bool isleft = o < len && (rks[id[k]] < rmid);
    bl[k] = __ballot_sync(0xffffffffu, isleft);
    wl += __popc(bl[k]);
  }
  if (lane == 0) wtot[w] = wl;
  __syncthreads();
  uint32_t wbase = 0, total = 0;
#pragma unroll
  for (int k = 0; k < LVL_THREADS / 32; ++k) {
    if (k < w) wbase += wtot[k];
    total += wtot[k];
  }
  if (w == 0) {  // publish this chunk's count, look back for the lefts of the earlier chunks, publish the prefix
    const unsigned long long epoch = ((unsigned long long)((lvl_ctl[LC_EPOCH] << 6) + (uint32_t)level + 1u) & 0x3fffffffull) << 34;
    unsigned long long* row = status + ((uint64_t)e * nseg + seg) * cps;
    uint32_t excl = 0;
    if (chunk > 0) {
      if (lane == 0) st_status(&row[chunk], epoch | (1ull << 32) | total);
      int j = (int)chunk - 1;
      uint32_t polls = 0;
      for (;;) {
        const int idx = j - lane;
        unsigned long long v = epoch | (2ull << 32);  // before chunk 0: an inclusive prefix of 0
        if (idx >= 0) v = ld_status(&row[idx]);
        const uint32_t state = ((v >> 34) == (epoch >> 34)) ? (uint32_t)(v >> 32) & 3u : 0u;
        const uint32_t have = __ballot_sync(0xffffffffu, state != 0u);
        const uint32_t incl = __ballot_sync(0xffffffffu, state == 2u);
        const int first = incl ? __ffs(incl) - 1 : 32;                      // nearest inclusive prefix in this window
        const uint32_t need = first >= 31 ? 0xffffffffu : ((2u << first) - 1u);  // lanes 0..first
        if ((have & need) == need) {
          uint32_t val = (lane <= first) ? (uint32_t)v : 0u;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
          excl += val;
          if (first < 32) break;
          j -= 32;
        } else if (++polls > (1u << 22)) {  // a predecessor never published (cannot happen): do not hang the GPU
          if (lane == 0) lvl_ctl[LC_ERR] = 1u;
          break;
        }
      }
    }
    if (lane == 0) {
      st_status(&row[chunk], epoch | (2ull << 32) | (unsigned long long)(excl + total));
      s_leftbase = excl;
    }
  }
  __syncthreads();
  const uint32_t leftbase = s_leftbase;
  uint32_t pre = wbase;
#pragma unroll
  for (int k = 0; k < IPT; ++k) {
    const uint32_t o = wbase_off + k * 32 + lane;
    if (o < len) {
      const uint32_t lrank = leftbase + pre + __popc(bl[k] & lt);  // lefts before me in the segment
      const bool isleft = (bl[k] >> lane) & 1u;
      const uint32_t dst = isleft ? a + lrank : mid + (o - lrank);
      lout[dst] = id[k];
    }
    pre += __popc(bl[k]);
  }
}

// ------------------------------------------------------------------------------------------ bottom levels

struct __align__(4) BotTab {
  uint16_t a, len, mid;
  uint8_t sd, kind;  // kind: 0 split, 1 leaf, 2 absent
  uint32_t node;
};
static_assert(sizeof(BotTab) == 12, "BotTab");

constexpr int BOT_THREADS = 512;
constexpr int BOT_IPT = BOT_CAP / BOT_THREADS;  // 4
constexpr int BOT_WARPS = BOT_THREADS / 32;

struct BotSmem {
  uint32_t gid[BOT_CAP];
  uint16_t lst[2][3][BOT_CAP];
  uint16_t segh[BOT_CAP];
  uint16_t scan[BOT_CAP];
  uint8_t side[BOT_CAP];
  uint32_t wtot[BOT_WARPS];
  int flag;
  BotTab tab[1];  // heap-indexed segment table, `heap` entries (dynamic shared memory; see bot_heap())
};

// Heap size of the bottom kernel's segment table: a segment of <= BOT_CAP particles is split while it holds more than
// mp, so the deepest local level is d = ceil(log2(BOT_CAP / mp)) and heap indices stay below 2^(d+1).  Sizing it to
// MAX_PARTS (6 KB instead of 24 KB at mp = 8) lets four CTAs share an SM, so the 512 segments of N = 1M run as one wave.
static inline uint32_t bot_heap(uint32_t mp) {
  uint32_t d = 0;
  while (((uint32_t)BOT_CAP >> d) > mp) ++d;
  return 2u << d;
}
static inline size_t bot_smem_bytes(uint32_t mp) { return sizeof(BotSmem) + (bot_heap(mp) - 1) * sizeof(BotTab); }

__global__ void __launch_bounds__(BOT_THREADS, 4)
build_bottom(Pos3c pos, const PosM* __restrict__ pm, Lists L, int level, uint32_t mp, int layout,
             const uint32_t* __restrict__ tstart, const uint32_t* __restrict__ tlen,
             const uint32_t* __restrict__ tnode, uint32_t* __restrict__ inv, WNode* __restrict__ nodes,
             double4* __restrict__ ms, uint32_t* __restrict__ perm, uint32_t* __restrict__ rank,
             PosM* __restrict__ posm, const uint32_t* __restrict__ flat, uint32_t heap) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BotSmem& S = *reinterpret_cast<BotSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t off = (1u << level) - 1;
  const uint32_t a0 = tstart[off + blockIdx.x], len0 = tlen[off + blockIdx.x], node0 = tnode[off + blockIdx.x];

  // ---- load: local slot j <-> id of the j-th entry of the x list
  for (uint32_t j = tid; j < len0; j += BOT_THREADS) {
    uint32_t g = L.l[0][a0 + j];
    S.gid[j] = g;
    inv[g] = j;
    S.lst[0][0][j] = (uint16_t)j;
    S.segh[j] = 1;
  }
  if (tid == 0) {
    BotTab t;
    t.a = 0;
    t.len = (uint16_t)len0;
    t.mid = 0;
    t.sd = 0;
    t.kind = 0;
    t.node = node0;
    S.tab[1] = t;
  }
  __syncthreads();
  const bool flat1 = flat[1] != 0, flat2 = flat[2] != 0;
  for (uint32_t j = tid; j < len0; j += BOT_THREADS) {
    if (!flat1) S.lst[0][1][j] = (uint16_t)inv[L.l[1][a0 + j]];
    if (!flat2) S.lst[0][2][j] = (uint16_t)inv[L.l[2][a0 + j]];
  }
  __syncthreads();

  int cur = 0, depth = 0;
  for (int lev = 0;; ++lev) {
    const uint32_t nn = 1u << lev;
    if (tid == 0) S.flag = 0;
    __syncthreads();
    // ---- node statistics for this local level
    for (uint32_t h = nn + tid; h < 2 * nn; h += BOT_THREADS) {
      BotTab t = S.tab[h];
      if (t.kind == 2) {
        if (2 * h + 1 < heap) S.tab[2 * h].kind = S.tab[2 * h + 1].kind = 2;
        continue;
      }
      if (t.len <= mp) {
        S.tab[h].kind = 1;
        if (2 * h + 1 < heap) S.tab[2 * h].kind = S.tab[2 * h + 1].kind = 2;
        continue;
      }
      double mn[3], mx[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        mn[d] = mx[d] = 0.0;
        if (d == 0 || !(d == 1 ? flat1 : flat2)) {
          mn[d] = pos.p[d][S.gid[S.lst[cur][d][t.a]]];
          mx[d] = pos.p[d][S.gid[S.lst[cur][d][t.a + t.len - 1]]];
        }
      }
      int sd = 0;
      double ext = mx[0] - mn[0];
#pragma unroll
      for (int d = 1; d < 3; ++d) {
        double e = mx[d] - mn[d];
        if (e > ext) {
          ext = e;
          sd = d;
        }
      }
      const uint32_t half = t.len / 2, mid = t.a + half;
      const double split_val = pos.p[sd][S.gid[S.lst[cur][sd][mid]]];
      const uint32_t nleft = (uint32_t)subtree_nodes(half, mp, layout);
      WNode* nd = &nodes[t.node];
      nd->size2 = __dmul_rn(ext, ext);
      nd->size = ext;
      nd->split_val = split_val;
      nd->a = t.node + 1 + nleft;
      nd->b = WN_INTERNAL | (uint32_t)sd;
      t.mid = (uint16_t)mid;
      t.sd = (uint8_t)sd;
      t.kind = 0;
      S.tab[h] = t;
      BotTab cl, cr;
      cl.a = t.a;
      cl.len = (uint16_t)half;
      cl.mid = 0;
      cl.sd = 0;
      cl.kind = 0;
      cl.node = t.node + 1;
      cr.a = (uint16_t)mid;
      cr.len = (uint16_t)(t.len - half);
      cr.mid = 0;
      cr.sd = 0;
      cr.kind = 0;
      cr.node = t.node + 1 + nleft;
      S.tab[2 * h] = cl;
      S.tab[2 * h + 1] = cr;
      S.flag = 1;
    }
    __syncthreads();
    depth = lev;
    if (!S.flag) break;
    // ---- side flags from the split-dimension list of each segment
    for (uint32_t i = tid; i < len0; i += BOT_THREADS) {
      BotTab t = S.tab[S.segh[i]];
      if (t.kind == 0) S.side[S.lst[cur][t.sd][i]] = (i >= t.mid) ? 1 : 0;
    }
    __syncthreads();
    // ---- stable partition of each list inside every segment
    for (int d = 0; d < 3; ++d) {
      if ((d == 1 && flat1) || (d == 2 && flat2)) continue;  // unused list
      uint32_t el[BOT_IPT], bl[BOT_IPT];
      uint32_t wl = 0;
#pragma unroll
      for (int k = 0; k < BOT_IPT; ++k) {
        uint32_t p = w * (32 * BOT_IPT) + k * 32 + lane;
        bool valid = p < len0;
        el[k] = valid ? S.lst[cur][d][p] : 0u;
        bool isleft = false;
        if (valid) {
          BotTab t = S.tab[S.segh[p]];
          isleft = (t.kind != 0) || (S.side[el[k]] == 0);
        }
        bl[k] = __ballot_sync(0xffffffffu, isleft);
        wl += __popc(bl[k]);
      }
      if (lane == 0) S.wtot[w] = wl;
      __syncthreads();
      uint32_t pre = 0;
      for (int k = 0; k < w; ++k) pre += S.wtot[k];
#pragma unroll
      for (int k = 0; k < BOT_IPT; ++k) {
        uint32_t p = w * (32 * BOT_IPT) + k * 32 + lane;
        if (p < len0) S.scan[p] = (uint16_t)(pre + __popc(bl[k] & lt));
        pre += __popc(bl[k]);
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BOT_IPT; ++k) {
        uint32_t p = w * (32 * BOT_IPT) + k * 32 + lane;
        if (p < len0) {
          BotTab t = S.tab[S.segh[p]];
          uint32_t dst = p;
          if (t.kind == 0) {
            uint32_t before = (uint32_t)S.scan[p] - (uint32_t)S.scan[t.a];  // lefts before me in my segment
            bool isleft = (bl[k] >> lane) & 1u;
            dst = isleft ? t.a + before : t.mid + (p - t.a - before);
          }
          S.lst[cur ^ 1][d][dst] = (uint16_t)el[k];
        }
      }
      __syncthreads();
    }
    // ---- descend: slot -> child segment
    for (uint32_t i = tid; i < len0; i += BOT_THREADS) {
      uint32_t h = S.segh[i];
      BotTab t = S.tab[h];
      if (t.kind == 0) S.segh[i] = (uint16_t)(2 * h + (i >= t.mid ? 1 : 0));
    }
    __syncthreads();
    cur ^= 1;
  }

  // ---- leaves.  (1) one thread per SLOT finds its place in the ascending particle-id order of its leaf (canonical
  //          leaf order) by counting the smaller ids among the <= MAX_PARTS members;
  //      (2) one thread per SLOT gathers {x,y,z,m} (one sector of the AoS copy, all gathers of the CTA in flight
  //          together) and writes perm / rank / posm coalesced;
  //      (3) one thread per leaf sums m and m*p sequentially in that order (array_kd_tree.rs:534-539 order of ops).
  const uint32_t hend = 2u << depth;
  uint16_t* order = S.lst[cur ^ 1][0];  // free buffer: tree slot (local) -> local id
  for (uint32_t p = tid; p < len0; p += BOT_THREADS) {
    const BotTab t = S.tab[S.segh[p]];  // the leaf that holds slot p
    const uint16_t l = S.lst[cur][0][p];
    const uint32_t v = S.gid[l];
    uint32_t r = 0;
    for (uint32_t k = 0; k < t.len; ++k) r += S.gid[S.lst[cur][0][t.a + k]] < v;
    order[t.a + r] = l;
  }
  __syncthreads();
  for (uint32_t p = tid; p < len0; p += BOT_THREADS) {
    const uint32_t id = S.gid[order[p]];
    const PosM q = pm[id];
    perm[a0 + p] = id;
    rank[id] = a0 + p;
    posm[a0 + p] = q;
  }
  __syncthreads();
  // {M, sum m*x, sum m*y, sum m*z} of the nodes of this segment live in shared memory (over the id / list buffers,
  // dead from here on); only the segment root's goes to global memory, for build_topup
  static_assert(sizeof(S.gid) + sizeof(S.lst) >= 1024 * 4 * sizeof(double), "node sums alias gid + lst (heap <= 1024)");
  double* msl = reinterpret_cast<double*>(smem_raw);
  for (uint32_t h = 1 + tid; h < hend; h += BOT_THREADS) {
    BotTab t = S.tab[h];
    if (t.kind != 1) continue;
    double m = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
    const uint32_t first = a0 + t.a;
    for (uint32_t k = 0; k < t.len; ++k) {
      const PosM q = posm[first + k];  // written by this CTA just above
      m = __dadd_rn(m, q.m);
      sx = __dadd_rn(sx, __dmul_rn(q.m, q.x));
      sy = __dadd_rn(sy, __dmul_rn(q.m, q.y));
      sz = __dadd_rn(sz, __dmul_rn(q.m, q.z));
    }
    msl[4 * h + 0] = m, msl[4 * h + 1] = sx, msl[4 * h + 2] = sy, msl[4 * h + 3] = sz;
    if (h == 1) ms[t.node] = make_double4(m, sx, sy, sz);
    WNode* nd = &nodes[t.node];
    nd->cx = sx;  // not part of the reference's Leaf; kept for debugging only
    nd->cy = sy;
    nd->cz = sz;
    nd->m = m;
    nd->size2 = 0.0;
    nd->size = 0.0;
    nd->split_val = 0.0;
    nd->a = first;
    nd->b = t.len;
  }
  __syncthreads();
  // ---- m / cm bottom-up inside this segment: internal = left + right, cm = sum / m (array_kd_tree.rs:548-550)
  for (int lev = depth - 1; lev >= 0; --lev) {
    const uint32_t nn = 1u << lev;
    for (uint32_t h = nn + tid; h < 2 * nn; h += BOT_THREADS) {
      BotTab t = S.tab[h];
      if (t.kind != 0) continue;
      const double* l = msl + 8 * h;  // children 2h and 2h + 1
      const double4 s = make_double4(__dadd_rn(l[0], l[4]), __dadd_rn(l[1], l[5]), __dadd_rn(l[2], l[6]), __dadd_rn(l[3], l[7]));
      msl[4 * h + 0] = s.x, msl[4 * h + 1] = s.y, msl[4 * h + 2] = s.z, msl[4 * h + 3] = s.w;
      if (h == 1) ms[t.node] = s;
      WNode* nd = &nodes[t.node];
      nd->m = s.x;
      nd->cx = __ddiv_rn(s.y, s.x);
      nd->cy = __ddiv_rn(s.z, s.x);
      nd->cz = __ddiv_rn(s.w, s.x);
    }
    __syncthreads();
  }
}

// m / cm for the global levels (single CTA; at most a few thousand nodes)
__global__ void __launch_bounds__(1024) build_topup(int l0, const uint32_t* __restrict__ tnode,
                                                    WNode* __restrict__ nodes, double4* __restrict__ ms) {
  pdl_sync();
  for (int lev = l0 - 1; lev >= 0; --lev) {
    const uint32_t nn = 1u << lev, off = nn - 1;
    for (uint32_t s = threadIdx.x; s < nn; s += blockDim.x) {
      const uint32_t node = tnode[off + s];
      WNode* nd = &nodes[node];
      const double4 l = ms[node + 1], r = ms[nd->a];
      double4 t = make_double4(__dadd_rn(l.x, r.x), __dadd_rn(l.y, r.y), __dadd_rn(l.z, r.z), __dadd_rn(l.w, r.w));
      ms[node] = t;
      nd->m = t.x;
      nd->cx = __ddiv_rn(t.y, t.x);
      nd->cy = __ddiv_rn(t.z, t.x);
      nd->cz = __ddiv_rn(t.w, t.x);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ host side

__global__ void fill_unused(WNode* nodes, uint64_t count) {
  pdl_sync();
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    WNode nd;
    nd.cx = nd.cy = nd.cz = nd.m = nd.size2 = nd.size = nd.split_val = 0.0;
    nd.a = 0xFFFFFFFFu;
    nd.b = WN_UNUSED;
    nodes[i] = nd;
  }
}

void init_unused_nodes(Ctx* c) {
  if (c->n_nodes == 0) return;
  KDNB_LAUNCH(c, fill_unused, (unsigned)((c->n_nodes + 255) / 256), 256, 0, c->nodes, c->n_nodes);
}

static bool g_bottom_attr_set = false;

int build_tree(Ctx* c) {
  const uint32_t n = (uint32_t)c->n;
  if (int rc = sort_lists(c)) return rc;
  if (!g_bottom_attr_set) {
    KDNB_CUDA_TRY(c, cudaFuncSetAttribute(build_bottom, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)bot_smem_bytes(4)));
    g_bottom_attr_set = true;
  }
  Pos3c pos = {{c->pos[0], c->pos[1], c->pos[2]}};
  KDNB_LAUNCH(c, build_root, 1, 64, 0, c->tstart, c->tlen, c->tnode, n, c->lvl_ctl);
  static const bool split_kernels = getenv("KDNB_LEVEL_SPLIT") != nullptr;  // profiling knob: count + scatter kernel pair
  int cur = 0;
  for (int lev = 0; lev < c->l0; ++lev) {
    const uint32_t nseg = 1u << lev;
    const uint32_t maxlen = (uint32_t)((c->n + nseg - 1) >> lev);
    const uint32_t cps = (maxlen + LVL_CHUNK - 1) / LVL_CHUNK;
    Lists Lin = {{c->list[cur], c->list[cur] + n, c->list[cur] + 2ull * n}};
    Lists Lout = {{c->list[cur ^ 1], c->list[cur ^ 1] + n, c->list[cur ^ 1] + 2ull * n}};
    if (split_kernels) {
      KDNB_LAUNCH(c, level_count, dim3(nseg * cps, 3), LVL_THREADS, 0, pos, Lin, lev, cps, c->mp, c->layout, c->tstart,
                  c->tlen, c->tnode, c->tmid, c->tsd, c->nodes, c->rk, n, c->tmr, c->chunk_cnt, c->flat);
      KDNB_LAUNCH(c, level_scatter, dim3(nseg * cps, 3), LVL_THREADS, 0, Lin, Lout, lev, cps, c->tstart, c->tlen,
                  c->tmid, c->tsd, c->rk, n, c->tmr, c->chunk_cnt, c->flat);
    } else {
      KDNB_LAUNCH(c, level_partition, nseg * cps * 3, LVL_THREADS, 0, pos, Lin, Lout, lev, cps, c->mp, c->layout,
                  c->tstart, c->tlen, c->tnode, c->tmid, c->tsd, c->nodes, c->rk, n, c->tmr,
                  reinterpret_cast<unsigned long long*>(c->lvl_status), c->lvl_ctl, c->flat);
    }
    cur ^= 1;
  }
  Lists Lb = {{c->list[cur], c->list[cur] + n, c->list[cur] + 2ull * n}};
  KDNB_LAUNCH(c, build_bottom, 1u << c->l0, BOT_THREADS, bot_smem_bytes(c->mp), pos, c->pm, Lb, c->l0, c->mp,
              c->layout, c->tstart, c->tlen, c->tnode, c->inv, c->nodes, c->ms, c->perm, c->rank, c->posm, c->flat,
              bot_heap(c->mp));
  if (c->l0 > 0) KDNB_LAUNCH(c, build_topup, 1, 1024, 0, c->l0, c->tnode, c->nodes, c->ms);
  KDNB_CHECK_LAUNCH(c);
  c->tree_valid = true;
  c->map_valid = true;
  return 0;
}

// ---- expand device nodes to the C-ABI record (kdnb_node), the mirror of `enum KDTree` (array_kd_tree.rs:18-34)
__global__ void export_nodes(const WNode* __restrict__ nodes, uint64_t count, kdnb_node* __restrict__ out) {
  pdl_sync();
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const WNode nd = nodes[i];
  kdnb_node o;
  o.split_dim = 0;
  o.num_parts = 0;
  o.leaf_first = KDNB_NO_INDEX;
  o.split_val = 0.0;
  o.m = 0.0;
  o.cm[0] = o.cm[1] = o.cm[2] = 0.0;
  o.size = 0.0;
  o.left = 0;
  o.right = 0;
  if (nd.b & WN_INTERNAL) {
    o.kind = KDNB_INTERNAL;
    o.split_dim = nd.b & 3u;
    o.split_val = nd.split_val;
    o.m = nd.m;
    o.cm[0] = nd.cx;
    o.cm[1] = nd.cy;
    o.cm[2] = nd.cz;
    o.size = nd.size;
    o.left = i + 1;
    o.right = nd.a;
  } else if (nd.b & WN_UNUSED) {
    o.kind = KDNB_LEAF;
  } else {
    o.kind = KDNB_LEAF;
    o.num_parts = nd.b;
    o.leaf_first = nd.a;
  }
  out[i] = o;
}

int export_tree(Ctx* c, kdnb_node* dev_out) {
  KDNB_LAUNCH(c, export_nodes, (unsigned)((c->n_nodes + 255) / 256), 256, 0, c->nodes, c->n_nodes, dev_out);
  KDNB_CHECK_LAUNCH(c);
  return 0;
}

}  // namespace kdnb
