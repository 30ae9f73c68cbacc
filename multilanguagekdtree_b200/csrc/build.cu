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
//  * levels whose segments are larger than BOT_CAP run as ONE global kernel each (level_partition);
//  * once segments fit (<= BOT_CAP slots) ONE kernel finishes all remaining levels in shared memory, emits the
//    leaves, the tree-ordered particle copies and sums m / sum(m*p) bottom-up inside the segment;
//  * a last single-CTA kernel carries m / sum(m*p) up the few global levels.
// m and cm are summed in the canonical order (leaf: ascending id, internal: left + right), see DESIGN.md.
#include <cstdlib>

#include "ctx.cuh"

namespace kdnb {

// the two ping-pong buffers of every per-dimension list: l[buffer][dimension]
struct Lists {
  uint32_t* l[2][3];
};
struct Pos3c {
  const double* p[3];
};

// ------------------------------------------------------------------------------------------ global levels

// Level table (device): one 16-byte record per segment, level l at offset 2^l - 1:
//   x = first slot, y = length, z = node index, w = buffer bits (bit d set: list d of this segment lives in buffer 1).
// The list of the split dimension is already partitioned (its first half IS the left child's list), so it stays where
// it is; only the other lists are partitioned into the other buffer, and the children inherit the flipped bits.
// lvl_ctl (device): [0..63] per-level CTA tickets, [64] build epoch, [65] look-back timeout flag
constexpr int LC_EPOCH = 64, LC_ERR = 65, LC_WORDS = 72;

__global__ void build_root(uint4* tseg, uint32_t n, uint32_t* lvl_ctl) {
  pdl_sync();
  if (threadIdx.x == 0) {
    tseg[0] = make_uint4(0u, n, 0u, 0u);  // the sort leaves every list in buffer 0
    lvl_ctl[LC_EPOCH] += 1u;
  }
  if (threadIdx.x < 64) lvl_ctl[threadIdx.x] = 0u;
}

// Statistics of segment s of `level` (bbox from the list ends, split dimension, median).  Every CTA working on the
// segment evaluates this (a handful of gathers); the one flagged `writer` also emits the node record and the table
// entries of the two children.  Returns sd, mid and the initial rank of the median element.
struct SegStats {
  uint32_t a, len, sd, mid, rmid, par;
};
__device__ __forceinline__ SegStats seg_stats(Pos3c pos, Lists L, int level, uint32_t s, uint32_t mp, int layout,
                                              uint4* __restrict__ tseg, WNode* __restrict__ nodes,
                                              const uint32_t* __restrict__ flat, const uint32_t* __restrict__ rk,
                                              uint32_t n, bool writer) {
  const uint32_t nseg = 1u << level;
  const uint32_t off = nseg - 1;
  const uint4 tb = tseg[off + s];
  const uint32_t a = tb.x, len = tb.y, node = tb.z, par = tb.w;
  const uint32_t half = len / 2, mid = a + half;
  // Three round trips after the table entry instead of five: the list ends AND the median entry of every dimension are
  // fetched together, then their coordinates and the median's initial rank; the split dimension only selects among
  // values that are already here (this chain is on the critical path of every CTA of the level).
  uint32_t ilo[3], ihi[3], imid[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    ilo[d] = ihi[d] = imid[d] = 0u;
    if (!flat[d]) {
      const uint32_t* lst = L.l[(par >> d) & 1u][d];
      ilo[d] = lst[a];
      ihi[d] = lst[a + len - 1];
      imid[d] = lst[mid];
    }
  }
  double mn[3], mx[3], vmid[3];
  uint32_t rmd[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    mn[d] = mx[d] = vmid[d] = 0.0;  // a flat dimension has extent 0 everywhere (its list is not maintained)
    rmd[d] = 0u;
    if (!flat[d]) {
      mn[d] = pos.p[d][ilo[d]];
      mx[d] = pos.p[d][ihi[d]];
      vmid[d] = pos.p[d][imid[d]];
      rmd[d] = rk[(uint64_t)d * n + imid[d]];
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
  SegStats r;
  r.a = a;
  r.len = len;
  r.sd = (uint32_t)sd;
  r.mid = mid;
  r.rmid = sd == 0 ? rmd[0] : (sd == 1 ? rmd[1] : rmd[2]);
  r.par = par;
  if (writer) {
    const double split_val = sd == 0 ? vmid[0] : (sd == 1 ? vmid[1] : vmid[2]);
    const uint32_t nleft = (uint32_t)subtree_nodes(half, mp, layout);
    WNode* nd = &nodes[node];
    nd->size2 = __dmul_rn(ext, ext);
    nd->size = ext;
    nd->split_val = split_val;
    nd->a = node + 1 + nleft;
    nd->b = WN_INTERNAL | (uint32_t)sd;
    const uint32_t cpar = par ^ (7u & ~(1u << sd));  // every list but the split dimension's changes buffer
    const uint32_t coff = 2 * nseg - 1;
    tseg[coff + 2 * s] = make_uint4(a, half, node + 1, cpar);
    tseg[coff + 2 * s + 1] = make_uint4(mid, len - half, node + 1 + nleft, cpar);
  }
  return r;
}

// ---- one kernel per global level: stable partition of every list (but the split dimension's) inside every segment,
// single pass.  "Goes left" is decided without any per-level flag pass: rk[d][id] is the rank of particle id in the
// INITIAL sorted list of dimension d (sort.cu).  Stable partitions keep every segment of list d ordered by rk[d], so the
// left half of a node split along sd is exactly { id : rk[sd][id] < rk[sd][id of the element at mid] }.
// A CTA owns one chunk (LVL_CHUNK slots) of one segment, for every list that has to move (one list for planar inputs,
// two otherwise).  It takes a ticket (CTAs are numbered in the order they START, so a CTA only ever waits for CTAs
// that are already running), evaluates the segment statistics, flags its entries and obtains the number of lefts in
// the earlier chunks of its segment by decoupled look-back over per-chunk status words
//     epoch (30 bits) | state (2 bits: 1 = chunk count, 2 = inclusive prefix) | value (32 bits)
// written and polled as single 64-bit words.  The epoch (build counter * 64 + level + 1) makes stale words of earlier
// launches unreadable, so the array is never cleared.  (History: a count + scatter kernel pair per level; then one
// look-back kernel with a CTA per (list, chunk) that copied the split-dimension list — two thirds of those CTAs only
// learnt that they had nothing to partition.)
template <int MINB>
__global__ void __launch_bounds__(LVL_THREADS, MINB)
level_partition(Pos3c pos, Lists L, int level, uint32_t cps, uint32_t mp, int layout, uint4* __restrict__ tseg,
                WNode* __restrict__ nodes, const uint32_t* __restrict__ rk, uint32_t n,
                unsigned long long* __restrict__ status, uint32_t* __restrict__ lvl_ctl,
                const uint32_t* __restrict__ flat, uint32_t seg_first) {
  pdl_sync();
  constexpr int IPT = LVL_CHUNK / LVL_THREADS;  // 8
  __shared__ uint32_t wtot[LVL_THREADS / 32];
  __shared__ SegStats st;
  __shared__ uint32_t s_ticket, s_leftbase;
  if (threadIdx.x == 0) s_ticket = atomicAdd(&lvl_ctl[level], 1u);
  __syncthreads();
  const uint32_t nseg = 1u << level;
  const uint32_t seg = seg_first + s_ticket / cps, chunk = s_ticket % cps;  // (seg_first > 0: this rank's subtree only)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  if (threadIdx.x == 0) st = seg_stats(pos, L, level, seg, mp, layout, tseg, nodes, flat, rk, n, chunk == 0);
  __syncthreads();
  const uint32_t a = st.a, len = st.len, sd = st.sd, rmid = st.rmid, mid = st.mid, par = st.par;
  if (chunk * (uint32_t)LVL_CHUNK >= len) return;  // (segments of a level differ by at most one slot; nobody looks back at these)
  const uint32_t* rks = rk + (uint64_t)sd * n;
  const uint32_t wbase_off = chunk * LVL_CHUNK + w * (32 * IPT);
  const unsigned long long epoch = ((unsigned long long)((lvl_ctl[LC_EPOCH] << 6) + (uint32_t)level + 1u) & 0x3fffffffull) << 34;
  for (uint32_t e = 0; e < 3; ++e) {
    if (e == sd || flat[e]) continue;  // the split-dimension list is already partitioned and stays in its buffer
    const uint32_t buf = (par >> e) & 1u;
    const uint32_t* lin = L.l[buf][e];
    uint32_t* lout = L.l[buf ^ 1u][e];
    uint32_t id[IPT], bl[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const uint32_t o = wbase_off + k * 32 + lane;
      id[k] = o < len ? lin[a + o] : 0u;
    }
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
}

// ------------------------------------------------------------------------------------------ bottom levels

struct __align__(4) BotTab {
  uint16_t a, len, mid;
  uint8_t sd, kind;  // kind: 0 split, 1 leaf, 2 absent
  uint32_t node;
};
static_assert(sizeof(BotTab) == 12, "BotTab");

#ifndef KDNB_BOT_THREADS
#define KDNB_BOT_THREADS 256
#endif
constexpr int BOT_THREADS = KDNB_BOT_THREADS;
constexpr int BOT_IPT = (BOT_CAP + BOT_THREADS - 1) / BOT_THREADS;  // (the last warp runs past BOT_CAP when this does not divide: every use is guarded by p < len0)
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

#ifndef KDNB_BOT_MINB
#define KDNB_BOT_MINB 6
#endif
__global__ void __launch_bounds__(BOT_THREADS, KDNB_BOT_MINB)
build_bottom(Pos3c pos, const PosM* __restrict__ pm, Lists L, int level, uint32_t mp, int layout,
             const uint4* __restrict__ tseg, uint32_t* __restrict__ inv, WNode* __restrict__ nodes,
             double4* __restrict__ ms, uint32_t* __restrict__ perm, uint32_t* __restrict__ rank,
             PosM* __restrict__ posm, const uint32_t* __restrict__ flat, uint32_t heap, uint32_t seg_first) {
  pdl_sync();
  KDNB_DYN_SMEM(smem_raw);
  BotSmem& S = *reinterpret_cast<BotSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t off = (1u << level) - 1;
  const uint4 tb = tseg[off + seg_first + blockIdx.x];
  const uint32_t a0 = tb.x, len0 = tb.y, node0 = tb.z;
  const uint32_t* const lx = L.l[tb.w & 1u][0];  // the buffer each list of this segment ended up in
  const uint32_t* const ly = L.l[(tb.w >> 1) & 1u][1];
  const uint32_t* const lz = L.l[(tb.w >> 2) & 1u][2];

  // ---- load: local slot j <-> id of the j-th entry of the x list
  for (uint32_t j = tid; j < len0; j += BOT_THREADS) {
    uint32_t g = lx[a0 + j];
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
    if (!flat1) S.lst[0][1][j] = (uint16_t)inv[ly[a0 + j]];
    if (!flat2) S.lst[0][2][j] = (uint16_t)inv[lz[a0 + j]];
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
      const uint32_t half = t.len / 2, mid = t.a + half;
      double mn[3], mx[3], vmid[3];  // the median coordinate of every dimension travels with the ends: one round trip
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        mn[d] = mx[d] = vmid[d] = 0.0;
        if (d == 0 || !(d == 1 ? flat1 : flat2)) {
          mn[d] = pos.p[d][S.gid[S.lst[cur][d][t.a]]];
          mx[d] = pos.p[d][S.gid[S.lst[cur][d][t.a + t.len - 1]]];
          vmid[d] = pos.p[d][S.gid[S.lst[cur][d][mid]]];
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
      const double split_val = sd == 0 ? vmid[0] : (sd == 1 ? vmid[1] : vmid[2]);
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
  static_assert(sizeof(S.gid) + sizeof(S.lst) >= (BOT_CAP / 2) * 4 * sizeof(double), "node sums alias gid + lst (heap <= BOT_CAP / 2 at MAX_PARTS >= 4)");
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

// m / cm for the global levels lev_hi-1 .. lev_lo (single CTA; at most a few thousand nodes).  Sharded builds run it
// twice: first over the levels below the shard level for this rank's subtree only (seg_shift = level of the shard,
// rank = its segment there), then — once every subtree root's sums have arrived — over the replicated top levels.
__global__ void __launch_bounds__(1024) build_topup(int lev_hi, int lev_lo, int shard_level, uint32_t shard_seg,
                                                    const uint4* __restrict__ tseg, WNode* __restrict__ nodes,
                                                    double4* __restrict__ ms) {
  pdl_sync();
  for (int lev = lev_hi - 1; lev >= lev_lo; --lev) {
    uint32_t nn = 1u << lev, first = 0;
    const uint32_t off = nn - 1;
    if (shard_level >= 0) nn = 1u << (lev - shard_level), first = shard_seg << (lev - shard_level);
    for (uint32_t s = first + threadIdx.x; s < first + nn; s += blockDim.x) {
      const uint32_t node = tseg[off + s].z;
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

// ------------------------------------------------------------------------------------------ sharded build: exchange
// Rank r has built the subtree of level-k segment r: node records [node_first, node_first + node_cnt) (a subtree is a
// contiguous index range in both layouts), tree slots [slot_first, slot_first + slot_cnt) of perm, and the root's mass
// sums.  Every thread loads a piece once and stores it into every peer's copy over NVLink; a system-scope fence and the
// last CTA then raise this rank's build flag on every peer (the same protocol as the accelerations, walk2.cuh).
__global__ void __launch_bounds__(256) subtree_push_kernel(P2PBuild pb, uint64_t node_first, uint64_t node_cnt,
                                                           uint32_t slot_first, uint32_t slot_cnt) {
  pdl_sync();
  const int me = pb.rank;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
  {
    const uint4* src = reinterpret_cast<const uint4*>(pb.nodes[me] + node_first);
    const uint64_t units = node_cnt * (sizeof(WNode) / sizeof(uint4));
    const uint32_t* words = reinterpret_cast<const uint32_t*>(src);
    for (uint64_t i = tid; i < units; i += nthr) {
      // slots the build never writes (about half of the padded layout) hold the default leaf on every rank already
      if (words[(i >> 2) * 16 + 11] & WN_UNUSED) continue;  // WNode::b
      const uint4 v = src[i];
      for (int p = 0; p < pb.world; ++p)
        if (p != me) reinterpret_cast<uint4*>(pb.nodes[p] + node_first)[i] = v;
    }
  }
  {
    const uint32_t* src = pb.perm[me] + slot_first;
    for (uint64_t i = tid; i < slot_cnt; i += nthr) {
      const uint32_t v = src[i];
      for (int p = 0; p < pb.world; ++p)
        if (p != me) pb.perm[p][slot_first + i] = v;
    }
  }
  if (tid == 0) {
    const double4 v = pb.ms[me][node_first];
    for (int p = 0; p < pb.world; ++p)
      if (p != me) pb.ms[p][node_first] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* st = pb.state[me];
    const uint32_t done = atomicAdd(&st[P2P_BDONE], 1u);
    if (done == gridDim.x - 1) {
      st[P2P_BDONE] = 0;
      __threadfence_system();
      const uint32_t epoch = st[P2P_BEPOCH];
      for (int p = 0; p < pb.world; ++p) *reinterpret_cast<volatile uint32_t*>(pb.state[p] + P2P_BFLAGS + me) = epoch + 1u;
    }
  }
}

// one CTA per GPU: wait until every rank (this one included) has published its subtree of the current build
__global__ void subtree_wait_kernel(uint32_t* state, int world) {
  pdl_sync();
  const uint32_t target = state[P2P_BEPOCH] + 1u;
  volatile uint32_t* flags = state + P2P_BFLAGS;
  if ((int)threadIdx.x < world) {
    const long long t0 = clock64();
    while (flags[threadIdx.x] < target) {
      if (clock64() - t0 > (30LL << 30)) {  // ~15 s: a peer died; record it instead of hanging for ever
        state[2] = 1u;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  if (threadIdx.x == 0) state[P2P_BEPOCH] = target;
}

// the tree-ordered particle copies and the id -> slot map of the slots OTHER ranks built: both follow from perm and the
// (replicated) particle state, so they are gathered here instead of travelling over NVLink (32 + 4 bytes per particle)
__global__ void __launch_bounds__(256) finish_foreign_kernel(uint32_t n, uint32_t own_first, uint32_t own_cnt,
                                                             const uint32_t* __restrict__ perm,
                                                             const PosM* __restrict__ pm, PosM* __restrict__ posm,
                                                             uint32_t* __restrict__ rank) {
  pdl_sync();
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n || (s - own_first) < own_cnt) return;
  const uint32_t id = perm[s];
  posm[s] = pm[id];
  rank[id] = s;
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

// level-k segment `seg` of the tree over n particles: first slot, length and node index (splits are always at len / 2,
// :560, and a left subtree occupies the node indices right after its parent, :566-581 / :120-126), host side
void subtree_of(uint64_t n, uint32_t mp, int layout, int k, uint32_t seg, uint64_t* a, uint64_t* len, uint64_t* node) {
  uint64_t fa = 0, fl = n, fn = 0;
  for (int lev = k - 1; lev >= 0; --lev) {
    const uint64_t left = fl / 2;
    if ((seg >> lev) & 1u) {
      fn += 1 + subtree_nodes(left, mp, layout);
      fa += left;
      fl -= left;
    } else {
      fn += 1;
      fl = left;
    }
  }
  *a = fa, *len = fl, *node = fn;
}

// Sharded build (this step): peer mode, world = 2^k, enough global levels, and a particle count where the exchange
// (about 31 bytes per particle sent to every peer, four more launches) costs less than the part of the lower levels it
// saves.  Measured (profiles/r02_ab_shard_build.txt): build 3.98 -> 3.19 ms at N=10M on 8 GPUs (step 7.12 -> 6.34 ms),
// 0.504 -> 0.468 ms at N=1M on 8 GPUs, 0.497 -> 0.459 / 3.98 -> 3.50 ms at N=1M / 10M on 2 GPUs; every rank's tree stays
// bit-identical to the single-GPU build (tests/multigpu_check.py).  KDNB_SHARD_BUILD=0 disables, =1 forces it wherever
// it is possible.
static int shard_level(const Ctx* c) {
  static const int mode = [] {
    const char* e = getenv("KDNB_SHARD_BUILD");
    return e ? (atoi(e) != 0 ? 1 : 0) : -1;
  }();
  if (mode == 0 || c->world <= 1 || !c->p2p_on) return 0;
  int k = 0;
  while ((1 << k) < c->world) ++k;
  if ((1 << k) != c->world || k > c->l0) return 0;
  if (mode < 0 && c->n < 500000ull) return 0;
  return k;
}

int build_tree(Ctx* c) {
  const uint32_t n = (uint32_t)c->n;
  if (int rc = sort_lists(c)) return rc;
  if (!c->bottom_attr_set) {  // per context: function attributes belong to the device the context drives
    KDNB_CUDA_TRY(c, cudaFuncSetAttribute(build_bottom, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)bot_smem_bytes(4)));
    c->bottom_attr_set = true;
  }
  Pos3c pos = {{c->pos[0], c->pos[1], c->pos[2]}};
  Lists L = {{{c->list[0], c->list[0] + n, c->list[0] + 2ull * n}, {c->list[1], c->list[1] + n, c->list[1] + 2ull * n}}};
  KDNB_LAUNCH(c, build_root, 1, 64, 0, c->tseg, n, c->lvl_ctl);
  // Resident CTAs per SM the partition kernel's register budget is sized for: 4 (63 registers, no spills) while a
  // level fits the GPU in about a wave, 8 (32 registers, 76 bytes spilled) when a level is many waves deep and the extra
  // CTAs hide the look-back and gather latencies (N = 10M: build 4.88 -> 4.62 ms; N = 1M: no difference).
  // KDNB_LVL_MINB overrides (profiling knob).
  static const int lvl_minb_env = [] {
    const char* e = getenv("KDNB_LVL_MINB");
    return e ? atoi(e) : 0;
  }();
  const int lvl_minb = lvl_minb_env ? lvl_minb_env : ((c->n + LVL_CHUNK - 1) / LVL_CHUNK > 8ull * c->num_sms ? 8 : 4);
  // multi-GPU: below level k every rank continues with the subtree of its own level-k segment only
  const int k = shard_level(c);
  c->shard_k = k;
  const uint32_t me = (uint32_t)c->rank_id;
  for (int lev = 0; lev < c->l0; ++lev) {
    const bool own = k > 0 && lev >= k;
    const uint32_t nseg = own ? 1u << (lev - k) : 1u << lev;
    const uint32_t seg_first = own ? me << (lev - k) : 0u;
    const uint32_t maxlen = (uint32_t)((c->n + (1ull << lev) - 1) >> lev);
    const uint32_t cps = (maxlen + LVL_CHUNK - 1) / LVL_CHUNK;
#define KDNB_LVL_ARGS pos, L, lev, cps, c->mp, c->layout, c->tseg, c->nodes, c->rk, n, \
                      reinterpret_cast<unsigned long long*>(c->lvl_status), c->lvl_ctl, c->flat, seg_first
    if (lvl_minb == 8) KDNB_LAUNCH(c, (level_partition<8>), nseg * cps, LVL_THREADS, 0, KDNB_LVL_ARGS);
    else if (lvl_minb == 6) KDNB_LAUNCH(c, (level_partition<6>), nseg * cps, LVL_THREADS, 0, KDNB_LVL_ARGS);
    else KDNB_LAUNCH(c, (level_partition<4>), nseg * cps, LVL_THREADS, 0, KDNB_LVL_ARGS);
#undef KDNB_LVL_ARGS
  }
  if (k == 0) {
    KDNB_LAUNCH(c, build_bottom, 1u << c->l0, BOT_THREADS, bot_smem_bytes(c->mp), pos, c->pm, L, c->l0, c->mp, c->layout,
                c->tseg, c->inv, c->nodes, c->ms, c->perm, c->rank, c->posm, c->flat, bot_heap(c->mp), 0u);
    if (c->l0 > 0) KDNB_LAUNCH(c, build_topup, 1, 1024, 0, c->l0, 0, -1, 0u, c->tseg, c->nodes, c->ms);
  } else {
    KDNB_LAUNCH(c, build_bottom, 1u << (c->l0 - k), BOT_THREADS, bot_smem_bytes(c->mp), pos, c->pm, L, c->l0, c->mp,
                c->layout, c->tseg, c->inv, c->nodes, c->ms, c->perm, c->rank, c->posm, c->flat, bot_heap(c->mp),
                me << (c->l0 - k));
    if (c->l0 > k) KDNB_LAUNCH(c, build_topup, 1, 1024, 0, c->l0, k, k, me, c->tseg, c->nodes, c->ms);
    uint64_t a = 0, len = 0, node = 0;
    subtree_of(c->n, c->mp, c->layout, k, me, &a, &len, &node);
    const uint64_t cnt = subtree_nodes(len, c->mp, c->layout);
    KDNB_LAUNCH(c, subtree_push_kernel, 4 * c->num_sms, 256, 0, c->p2pb, node, cnt, (uint32_t)a, (uint32_t)len);
    KDNB_LAUNCH(c, subtree_wait_kernel, 1, 32, 0, c->p2p_state, c->world);
    KDNB_LAUNCH(c, finish_foreign_kernel, (n + 255) / 256, 256, 0, n, (uint32_t)a, (uint32_t)len, c->perm, c->pm, c->posm,
                c->rank);
    KDNB_LAUNCH(c, build_topup, 1, 1024, 0, k, 0, -1, 0u, c->tseg, c->nodes, c->ms);
  }
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
