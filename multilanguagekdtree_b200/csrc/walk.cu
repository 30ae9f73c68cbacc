// walk.cu — theta-criterion force walk, calc_accel / accel_recur of the reference
// (Parallel/RustVersion/src/array_kd_tree.rs:585-621) with calc_pp_accel (array_particle.rs:67-76) in the leaves.
//
// Mapping: one warp = 32*PPL consecutive TREE-ORDERED particles (spatially compact: 4-7 adjacent leaves per 32),
// PPL particles per lane.  The warp walks the tree once with a shared-memory stack of (node, lane masks) entries:
//   * every lane in the entry's mask evaluates the reference's acceptance test for ITS OWN particle(s),
//       size*size < (THETA*THETA) * dist_sqr            (array_kd_tree.rs:606)
//     with the same unfused operation order, so each particle accepts / opens exactly the nodes the
//     reference does (checked by the per-particle visit counters, KDNB_FLAG_WALK_COUNTS);
//   * an accepted node is NOT evaluated on the spot: its monopole {cm, m} and the masks of accepting lanes are
//     appended to a per-warp interaction list in shared memory; if any lane still has to open the node
//     (__ballot_sync) both children are pushed with the remaining masks;
//   * a leaf appends its particles (loaded by up to MAX_PARTS lanes in one coalesced access) with the masks of
//     lanes that reached it, minus the lane that owns the particle (leaf_parts[i] != p, :590);
//   * when the list is full it is drained: all lanes stream over the point masses (shared-memory broadcast loads,
//     no global loads, no divergent control flow) and accumulate -m * d / r^3 under their mask bit.
// Traversal (latency-bound pointer chasing, 10 FP64 ops per test) and force evaluation (FP64-pipe-bound, 16 ops
// per interaction, unrolled) are thereby decoupled.  Node records are one 64-byte line; every node load is
// warp-uniform.
// Accumulation is a running f64 sum per particle (the reference combines pairwise along the recursion, :611-613;
// the difference is summation order only and is covered by the stated 1e-12 tolerance).
//
// Tried and rejected on the GPU (profiles/README.md): per-lane interaction queues with gathered record loads
// (L1-bound, 5.8 ms vs 4.7 ms at N=1M) and persistent CTAs with static contiguous ranges (tail imbalance).
#include <cstdlib>

#include "ctx.cuh"

namespace kdnb {

constexpr int WALK_THREADS = 128;
constexpr int WALK_WARPS = WALK_THREADS / 32;
constexpr int WALK_STACK = 40;  // deepest stack = tree depth + 2 (<= 27 at 1e8 particles)
constexpr int WALK_LIST = 64;   // interaction-list capacity per warp (>= 2 * largest MAX_PARTS)

struct __align__(32) Rec32 {
  double a, b, c, d;
};

// -m / r^3 without divide or sqrt: y0 = MUFU.RSQ64H estimate (rel. error < 2^-22), e = 1 - d2*y0^2,
// r^-3 = y0^3 * (1 - e)^(-3/2) = y0^3 * (1 + 1.5 e + 1.875 e^2 + O(e^3)); O(e^3) < 2^-63.  No special cases:
// callers discard the result by select when the pair is masked out (d2 == 0 gives NaN there).
__device__ __forceinline__ double neg_m_over_r3_fast(double mneg, double d2) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d2));
  const double y2 = __dmul_rn(y, y);
  const double e = fma(-d2, y2, 1.0);
  const double y3 = __dmul_rn(y, y2);
  const double q = fma(1.875, e, 1.5);
  const double mq = __dmul_rn(mneg, y3);
  return fma(__dmul_rn(mq, e), q, mq);
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int PPL>
struct WalkSmem {
  uint4 stk[WALK_WARPS][WALK_STACK];   // {node, mask0, mask1, -}
  Rec32 lpos[WALK_WARPS][WALK_LIST];   // {x, y, z, m} of a monopole or of a leaf particle
  uint4 lmask[WALK_WARPS][WALK_LIST];  // {mask0, mask1, is_particle, -}
};

template <bool EXACT>
__device__ __forceinline__ void interact(const Rec32& e, bool use, bool is_particle, double px, double py, double pz,
                                         double& ax, double& ay, double& az) {
  const double dx = __dsub_rn(px, e.a), dy = __dsub_rn(py, e.b), dz = __dsub_rn(pz, e.c);
  if (EXACT) {
    if (use) {
      const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      const double dist = __dsqrt_rn(d2);
      // node: -m / (dist_sqr * dist) (array_kd_tree.rs:608); particle: -m / (dist*dist*dist) (array_particle.rs:72)
      const double den = is_particle ? __dmul_rn(__dmul_rn(dist, dist), dist) : __dmul_rn(d2, dist);
      const double magi = __ddiv_rn(-e.d, den);
      ax = __dadd_rn(ax, __dmul_rn(magi, dx));
      ay = __dadd_rn(ay, __dmul_rn(magi, dy));
      az = __dadd_rn(az, __dmul_rn(magi, dz));
    }
  } else {
    const double d2 = fma(dz, dz, fma(dy, dy, __dmul_rn(dx, dx)));
    double magi = neg_m_over_r3_fast(-e.d, d2);
    magi = use ? magi : 0.0;
    ax = fma(magi, dx, ax);
    ay = fma(magi, dy, ay);
    az = fma(magi, dz, az);
  }
}

template <int PPL, bool EXACT, bool COUNTS>
__device__ __forceinline__ void drain_list(const Rec32* __restrict__ lpos, const uint4* __restrict__ lmask, int cnt,
                                           int lane, const double (&px)[PPL], const double (&py)[PPL],
                                           const double (&pz)[PPL], double (&ax)[PPL], double (&ay)[PPL],
                                           double (&az)[PPL], unsigned long long (&cp)[PPL]) {
#pragma unroll 2
  for (int i = 0; i < cnt; ++i) {
    const Rec32 e = lpos[i];
    const uint4 mk = lmask[i];
#pragma unroll
    for (int u = 0; u < PPL; ++u) {
      const uint32_t m = (u == 0 ? mk.x : mk.y);
      const bool use = (m >> lane) & 1u;
      interact<EXACT>(e, use, mk.z != 0, px[u], py[u], pz[u], ax[u], ay[u], az[u]);
      if (COUNTS) cp[u] += (use && mk.z) ? 1 : 0;
    }
  }
}

template <int PPL, int MINB, bool PF, bool EXACT, bool COUNTS>
__global__ void __launch_bounds__(WALK_THREADS, MINB)
walk_kernel(const WNode* __restrict__ nodes, const PosM* __restrict__ posm, double* __restrict__ acc_t,
            uint32_t slot_begin, uint32_t slot_end, double theta2, uint32_t max_parts,
            unsigned long long* __restrict__ wcounts) {
  __shared__ WalkSmem<PPL> S;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t base = slot_begin + (blockIdx.x * WALK_WARPS + w) * (32 * PPL);
  uint32_t slot[PPL], mask[PPL];
  double px[PPL], py[PPL], pz[PPL], ax[PPL], ay[PPL], az[PPL];
  unsigned long long cv[PPL], ca[PPL], cl[PPL], cp[PPL];
#pragma unroll
  for (int u = 0; u < PPL; ++u) {
    slot[u] = base + u * 32 + lane;
    const bool valid = slot[u] < slot_end;
    mask[u] = __ballot_sync(0xffffffffu, valid);
    px[u] = py[u] = pz[u] = 0.0;
    if (valid) {
      const PosM me = posm[slot[u]];
      px[u] = me.x;
      py[u] = me.y;
      pz[u] = me.z;
    }
    ax[u] = ay[u] = az[u] = 0.0;
    cv[u] = ca[u] = cl[u] = cp[u] = 0;
  }
  if (mask[0] == 0) return;

  uint4* st = S.stk[w];
  Rec32* lpos = S.lpos[w];
  uint4* lmask = S.lmask[w];
  int sp = 0, ln = 0;
  uint32_t node = 0;  // the root, with every valid lane in the masks
  bool more = true;
  while (more) {
    const Rec32* rec = reinterpret_cast<const Rec32*>(nodes + node);
    const Rec32 c = rec[0];                                           // cx, cy, cz, m
    const int4 info = __ldg(reinterpret_cast<const int4*>(rec + 1));  // size2, (a, b)
    if (PF) prefetch_l1(rec + 2);  // the left child is the next record
    const uint32_t na = (uint32_t)info.z, nb = (uint32_t)info.w;
    bool descend = false;
    bool in[PPL];
#pragma unroll
    for (int u = 0; u < PPL; ++u) in[u] = (mask[u] >> lane) & 1u;
    if (nb & WN_INTERNAL) {
      const double size2 = __hiloint2double(info.y, info.x);
      uint32_t am[PPL], any = 0, open_any = 0;
#pragma unroll
      for (int u = 0; u < PPL; ++u) {
        const double dx = __dsub_rn(px[u], c.a), dy = __dsub_rn(py[u], c.b), dz = __dsub_rn(pz[u], c.c);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));  // :604
        const bool accept = in[u] && (size2 < __dmul_rn(theta2, d2));                                     // :606
        am[u] = __ballot_sync(0xffffffffu, accept);
        any |= am[u];
        mask[u] &= ~am[u];
        open_any |= mask[u];
        if (COUNTS) {
          cv[u] += in[u];
          ca[u] += accept;
        }
      }
      if (any) {
        if (lane == 0) {
          lpos[ln] = c;
          lmask[ln] = make_uint4(am[0], PPL > 1 ? am[PPL - 1] : 0u, 0u, 0u);
        }
        ln += 1;
      }
      descend = open_any != 0;
      if (descend) {
        st[sp++] = make_uint4(na, mask[0], PPL > 1 ? mask[PPL - 1] : 0u, 0u);  // right child waits on the stack
        if (PF) prefetch_l1(nodes + na);
        node = node + 1;  // left child first, as the recursion (:611): it is the next record, masks stay in registers
      }
    } else {
      const uint32_t cnt = nb;
      if (COUNTS) {
#pragma unroll
        for (int u = 0; u < PPL; ++u) cl[u] += in[u];
      }
      if ((uint32_t)lane < cnt) {
        const uint32_t j = na + lane;
        const PosM q = posm[j];
        Rec32 r;
        r.a = q.x;
        r.b = q.y;
        r.c = q.z;
        r.d = q.m;
        uint32_t mk[PPL];
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
          const uint32_t t = j - (base + u * 32);  // the lane that owns particle j, if it is one of ours
          mk[u] = t < 32u ? (mask[u] & ~(1u << t)) : mask[u];
        }
        lpos[ln + lane] = r;
        lmask[ln + lane] = make_uint4(mk[0], PPL > 1 ? mk[PPL - 1] : 0u, 1u, 0u);
      }
      ln += cnt;
    }
    if (ln + (int)max_parts > WALK_LIST) {
      __syncwarp();
      drain_list<PPL, EXACT, COUNTS>(lpos, lmask, ln, lane, px, py, pz, ax, ay, az, cp);
      __syncwarp();
      ln = 0;
    }
    if (!descend) {
      more = sp > 0;
      if (more) {
        const uint4 e = st[--sp];
        node = e.x;
        mask[0] = e.y;
        if (PPL > 1) mask[PPL - 1] = e.z;
      }
    }
  }
  __syncwarp();
  drain_list<PPL, EXACT, COUNTS>(lpos, lmask, ln, lane, px, py, pz, ax, ay, az, cp);
#pragma unroll
  for (int u = 0; u < PPL; ++u) {
    if (slot[u] < slot_end) {
      acc_t[3ull * slot[u] + 0] = ax[u];
      acc_t[3ull * slot[u] + 1] = ay[u];
      acc_t[3ull * slot[u] + 2] = az[u];
      if (COUNTS) {
        wcounts[4ull * slot[u] + 0] = cv[u];
        wcounts[4ull * slot[u] + 1] = ca[u];
        wcounts[4ull * slot[u] + 2] = cl[u];
        wcounts[4ull * slot[u] + 3] = cp[u];
      }
    }
  }
}

template <int PPL, int MINB, bool PF>
static void launch_walk(Ctx* c, uint32_t begin, uint32_t end) {
  const uint32_t groups = (end - begin + 32 * PPL - 1) / (32 * PPL);
  const uint32_t grid = (groups + WALK_WARPS - 1) / WALK_WARPS;
  const bool counts = (c->flags & KDNB_FLAG_WALK_COUNTS) != 0;
  const bool exact = (c->flags & KDNB_FLAG_EXACT_MATH) != 0;
#define KDNB_WALK_ARGS c->nodes, c->posm, c->acc_t, begin, end, c->theta2, c->mp, c->wcounts
  if (exact && counts)
    KDNB_LAUNCH(c, (walk_kernel<PPL, MINB, PF, true, true>), grid, WALK_THREADS, 0, KDNB_WALK_ARGS);
  else if (exact)
    KDNB_LAUNCH(c, (walk_kernel<PPL, MINB, PF, true, false>), grid, WALK_THREADS, 0, KDNB_WALK_ARGS);
  else if (counts)
    KDNB_LAUNCH(c, (walk_kernel<PPL, MINB, PF, false, true>), grid, WALK_THREADS, 0, KDNB_WALK_ARGS);
  else
    KDNB_LAUNCH(c, (walk_kernel<PPL, MINB, PF, false, false>), grid, WALK_THREADS, 0, KDNB_WALK_ARGS);
#undef KDNB_WALK_ARGS
}

int walk(Ctx* c) {
  const uint32_t n = (uint32_t)c->n;
  uint32_t begin = 0, end = n;
  if (c->world > 1) {
    begin = (uint32_t)std::min<uint64_t>(c->n, (uint64_t)c->rank_id * c->shard_slots);
    end = (uint32_t)std::min<uint64_t>(c->n, (uint64_t)(c->rank_id + 1) * c->shard_slots);
  }
  if (end > begin) {
    static const int cfg = [] {
      const char* s = getenv("KDNB_WALK_CFG");  // tuning knob for profiling runs: <ppl><min blocks per SM>
      return s ? atoi(s) : 0;
    }();
    switch (cfg) {
      case 14: launch_walk<1, 4, false>(c, begin, end); break;
      case 18: launch_walk<1, 8, false>(c, begin, end); break;
      case 110: launch_walk<1, 10, false>(c, begin, end); break;
      case 112: launch_walk<1, 12, false>(c, begin, end); break;
      case 24: launch_walk<2, 4, false>(c, begin, end); break;
      case 26: launch_walk<2, 6, false>(c, begin, end); break;
      case 27: launch_walk<2, 7, false>(c, begin, end); break;
      case 28: launch_walk<2, 8, false>(c, begin, end); break;
      case 210: launch_walk<2, 10, false>(c, begin, end); break;
      case 127: launch_walk<2, 7, true>(c, begin, end); break;
      case 118: launch_walk<1, 8, true>(c, begin, end); break;
      default: launch_walk<2, 7, true>(c, begin, end); break;
    }
    KDNB_CHECK_LAUNCH(c);
  }
  c->acc_valid = true;
  return 0;
}

}  // namespace kdnb
