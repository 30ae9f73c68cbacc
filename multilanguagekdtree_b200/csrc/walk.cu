// walk.cu — theta-criterion force walk, calc_accel / accel_recur of the reference
// (Parallel/RustVersion/src/array_kd_tree.rs:585-621) with calc_pp_accel (array_particle.rs:67-76) in the leaves.
//
// Mapping: one warp = 32 consecutive TREE-ORDERED particles (4-7 adjacent leaves, spatially compact), one lane
// per particle.  The warp walks the tree once with a shared-memory stack of (node, lane mask) entries:
//   * every lane in the entry's mask evaluates the reference's acceptance test for ITS OWN particle,
//       size*size < (THETA*THETA) * dist_sqr            (array_kd_tree.rs:606)
//     with the same unfused operation order, so each particle accepts / opens exactly the nodes the
//     reference does (checked by the per-particle visit counters, KDNB_FLAG_WALK_COUNTS);
//   * lanes that accept add the monopole and leave the mask; if any lane still has to open the node
//     (__ballot_sync) both children are pushed with the remaining mask;
//   * leaves: every lane still in the mask sums the direct pair forces, skipping itself (:590).
// Node records are one 64-byte line each and every load is warp-uniform (one wavefront, L1/L2 resident).
// Accumulation is a running f64 sum per lane (the reference combines pairwise along the recursion, :611-613;
// the difference is summation order only and is covered by the stated 1e-12 tolerance).
#include "ctx.cuh"

namespace kdnb {

constexpr int WALK_THREADS = 128;
constexpr int WALK_WARPS = WALK_THREADS / 32;
constexpr int WALK_STACK = 64;

// EXACT: sqrt + divide exactly as the reference writes them; otherwise rsqrt-based (<= 2 ulp apart).
template <bool EXACT>
__device__ __forceinline__ double inv_r3_times(double mneg, double d2) {
  if (EXACT) {
    double dist = __dsqrt_rn(d2);
    return __ddiv_rn(mneg, __dmul_rn(d2, dist));  // -m / (dist_sqr * dist), array_kd_tree.rs:608
  } else {
    double r = rsqrt(d2);
    return __dmul_rn(__dmul_rn(mneg, r), __dmul_rn(r, r));
  }
}

template <bool EXACT, bool COUNTS>
__global__ void __launch_bounds__(WALK_THREADS)
walk_kernel(const WNode* __restrict__ nodes, const PosM* __restrict__ posm, double* __restrict__ acc_t,
            uint32_t slot_begin, uint32_t slot_end, double theta2, unsigned long long* __restrict__ wcounts) {
  __shared__ uint2 stk[WALK_WARPS][WALK_STACK];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t slot = slot_begin + (blockIdx.x * WALK_WARPS + w) * 32 + lane;
  const bool valid = slot < slot_end;
  const uint32_t mask0 = __ballot_sync(0xffffffffu, valid);
  if (mask0 == 0) return;
  double px = 0.0, py = 0.0, pz = 0.0;
  if (valid) {
    const PosM me = posm[slot];
    px = me.x;
    py = me.y;
    pz = me.z;
  }
  double ax = 0.0, ay = 0.0, az = 0.0;
  unsigned long long cv = 0, ca = 0, cl = 0, cp = 0;

  uint2* st = stk[w];
  int sp = 0;
  st[sp++] = make_uint2(0u, mask0);
  while (sp > 0) {
    const uint2 e = st[--sp];
    const uint32_t node = e.x, mask = e.y;
    const bool in = (mask >> lane) & 1u;
    const int4 info = __ldg(reinterpret_cast<const int4*>(reinterpret_cast<const char*>(nodes + node) + 32));
    const uint32_t na = (uint32_t)info.z, nb = (uint32_t)info.w;
    if (nb & WN_INTERNAL) {
      const double size2 = __hiloint2double(info.y, info.x);
      const double2 c01 = __ldg(reinterpret_cast<const double2*>(nodes + node));
      const double2 c23 = __ldg(reinterpret_cast<const double2*>(nodes + node) + 1);
      const double dx = __dsub_rn(px, c01.x), dy = __dsub_rn(py, c01.y), dz = __dsub_rn(pz, c23.x);
      const double d2 =
          __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));  // :604, left-assoc, unfused
      const bool accept = in && (size2 < __dmul_rn(theta2, d2));                             // :606
      const uint32_t am = __ballot_sync(0xffffffffu, accept);
      if (accept) {
        const double magi = inv_r3_times<EXACT>(-c23.y, d2);  // :607-608
        if (EXACT) {
          ax = __dadd_rn(ax, __dmul_rn(dx, magi));
          ay = __dadd_rn(ay, __dmul_rn(dy, magi));
          az = __dadd_rn(az, __dmul_rn(dz, magi));
        } else {
          ax = fma(dx, magi, ax);
          ay = fma(dy, magi, ay);
          az = fma(dz, magi, az);
        }
      }
      if (COUNTS) {
        cv += in;
        ca += accept;
      }
      const uint32_t open = mask & ~am;
      if (open) {
        st[sp++] = make_uint2(na, open);        // right
        st[sp++] = make_uint2(node + 1, open);  // left is visited first, as the recursion does (:611)
      }
    } else {
      const uint32_t cnt = nb;
      if (COUNTS) cl += in;
      for (uint32_t k = 0; k < cnt; ++k) {
        const uint32_t j = na + k;
        const double2 q01 = __ldg(reinterpret_cast<const double2*>(posm + j));
        const double2 q23 = __ldg(reinterpret_cast<const double2*>(posm + j) + 1);
        if (in && j != slot) {  // leaf_parts[i] != p (:590); slots are a permutation of ids
          const double dx = __dsub_rn(px, q01.x), dy = __dsub_rn(py, q01.y), dz = __dsub_rn(pz, q23.x);
          if (EXACT) {
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            const double dist = __dsqrt_rn(d2);
            const double magi = __ddiv_rn(-q23.y, __dmul_rn(__dmul_rn(dist, dist), dist));  // array_particle.rs:72
            ax = __dadd_rn(ax, __dmul_rn(magi, dx));
            ay = __dadd_rn(ay, __dmul_rn(magi, dy));
            az = __dadd_rn(az, __dmul_rn(magi, dz));
          } else {
            const double d2 = fma(dz, dz, fma(dy, dy, __dmul_rn(dx, dx)));
            const double r = rsqrt(d2);
            const double magi = __dmul_rn(__dmul_rn(-q23.y, r), __dmul_rn(r, r));
            ax = fma(magi, dx, ax);
            ay = fma(magi, dy, ay);
            az = fma(magi, dz, az);
          }
          if (COUNTS) cp += 1;
        }
      }
    }
  }
  if (valid) {
    acc_t[3ull * slot + 0] = ax;
    acc_t[3ull * slot + 1] = ay;
    acc_t[3ull * slot + 2] = az;
    if (COUNTS) {
      wcounts[4ull * slot + 0] = cv;
      wcounts[4ull * slot + 1] = ca;
      wcounts[4ull * slot + 2] = cl;
      wcounts[4ull * slot + 3] = cp;
    }
  }
}

int walk(Ctx* c) {
  const uint32_t n = (uint32_t)c->n;
  uint32_t begin = 0, end = n;
  if (c->world > 1) {
    begin = (uint32_t)std::min<uint64_t>(c->n, (uint64_t)c->rank_id * c->shard_slots);
    end = (uint32_t)std::min<uint64_t>(c->n, (uint64_t)(c->rank_id + 1) * c->shard_slots);
  }
  if (end > begin) {
    const uint32_t warps = (end - begin + 31) / 32;
    const uint32_t grid = (warps + WALK_WARPS - 1) / WALK_WARPS;
    const bool counts = (c->flags & KDNB_FLAG_WALK_COUNTS) != 0;
    const bool exact = (c->flags & KDNB_FLAG_EXACT_MATH) != 0;
    if (exact && counts)
      KDNB_LAUNCH(c, (walk_kernel<true, true>), grid, WALK_THREADS, 0, c->nodes, c->posm, c->acc_t, begin, end, c->theta2, c->wcounts);
    else if (exact)
      KDNB_LAUNCH(c, (walk_kernel<true, false>), grid, WALK_THREADS, 0, c->nodes, c->posm, c->acc_t, begin, end, c->theta2, c->wcounts);
    else if (counts)
      KDNB_LAUNCH(c, (walk_kernel<false, true>), grid, WALK_THREADS, 0, c->nodes, c->posm, c->acc_t, begin, end, c->theta2, c->wcounts);
    else
      KDNB_LAUNCH(c, (walk_kernel<false, false>), grid, WALK_THREADS, 0, c->nodes, c->posm, c->acc_t, begin, end, c->theta2, c->wcounts);
    KDNB_CHECK_LAUNCH(c);
  }
  c->acc_valid = true;
  return 0;
}

}  // namespace kdnb
