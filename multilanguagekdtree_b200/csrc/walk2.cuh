// walk2.cuh — the walk kernel (launched from walk.cu).
//
// theta-criterion force walk, calc_accel / accel_recur of the reference
// (Parallel/RustVersion/src/array_kd_tree.rs:585-621) with calc_pp_accel (array_particle.rs:67-76) in the leaves.
//
// One warp (= one CTA, 24 or 28 CTAs per SM: walk.cu) owns 32 consecutive TREE-ORDERED particles (a compact patch of
// ~4 leaves), one per lane, and keeps a shared-memory stack of frontier entries (node, lane mask).  Per round it pops up to 32
// entries and classifies them ONE NODE PER LANE against the bounding box of the 32 particles:
//   far   : size^2 <  theta^2 * dmin^2 * (1 - 1e-9)  -> every particle in the entry's mask accepts the node
//   near  : size^2 >= theta^2 * dmax^2 * (1 + 1e-9)  -> every particle opens it
//   mixed : otherwise -> the reference's test is evaluated per particle (below)
//   leaf  : its particles become pair interactions
// dmin/dmax bound the distances from the node's centre of mass to the box from below / above (box_bounds()); the
// 1e-9 margin dwarfs the rounding of either side, so "far"/"near" provably agree with the reference's per-particle test
//       size*size < (THETA*THETA) * dist_sqr            (array_kd_tree.rs:606)
// Mixed nodes of a batch are parked in shared memory ({cm, size^2}, compacted); the warp then runs over them with ONE
// PARTICLE PER LANE evaluating exactly that test with the reference's unfused operation order (:601-606), and a ballot
// per node — left in shared memory for the owning lane — is the accept mask.  After that the batch is finished one node
// per lane again: far + accepted monopoles are appended to the interaction list with their lane masks, near + still-open
// nodes push both children.  Every particle thus accepts / opens precisely the nodes the reference's recursion does
// (checked by KDNB_FLAG_WALK_COUNTS against the oracle: per-particle counts of tests, accepts, leaf visits and pair
// interactions are identical).  Leaves of a batch are queued and expanded 32/LP leaves at a time (LP lanes per leaf):
// lane -> (leaf, k) loads leaf particle k with one 32-byte access and appends it with the mask of the lanes that
// reached the leaf minus the owner (leaf_parts[i] != p, :590).
//
// Why this shape: on B200 an FP64 warp instruction holds its scheduler's issue port for two cycles and nothing
// else issues in its shadow (tools/issue_probe.cu: 16 DFMA = 35 cycles, every extra ALU instruction +1 cycle), so the
// kernel is bound by  ~2 * FP64 instructions + other instructions  (profiles/README.md: the ncu instruction counts
// reproduce the measured time), and every change is judged by the instructions it removes.
//
// Forces are not evaluated during the traversal: the interaction list is drained by a branch-free loop over blocks of
// four interactions (W2Blk: one base register, seven 16-byte broadcast loads), two blocks per iteration, 13 FP64 + 6.3
// other instructions per interaction on planar inputs; an incomplete block waits for the next drain instead of being
// padded.  When every z coordinate is +-0 and every mass is > 0 (flat[3], sort.cu: sort_prep —
// the reference's own initial conditions, circular_orbits, array_particle.rs:19-44, are planar for ever) all
// centre-of-mass z are +-0 too, dz is exactly 0 in every interaction and every test, and the z terms are skipped: the
// results are bit-identical to the general path, az stays +0.
//
// Accumulation is a running f64 sum per particle (the reference combines pairwise along the recursion, :611-613);
// the difference is summation order only and is covered by the stated 1e-12 tolerance.
#pragma once
#include "ctx.cuh"

namespace kdnb {

#ifndef KDNB_W2_STACK
#define KDNB_W2_STACK 304
#endif
constexpr int W2_STACK = KDNB_W2_STACK;  // soft capacity: batches shrink as the stack fills (deepest use seen: 244 entries at N = 1M,
                                         // 286 at 10M, tests/devtools/walk_model.c); 304 + the layout below = 7296 bytes for the
                                         // production kernel, so that 28 CTAs fit an SM (walk.cu)
constexpr int W2_SLACK = 32;   // depth-first tail when the stack is at capacity (tree depth <= 27 at 1e8 particles)
#ifndef KDNB_W2_LIST
#define KDNB_W2_LIST 96
#endif
#ifndef KDNB_W2_LIST_FLAT
#define KDNB_W2_LIST_FLAT (KDNB_W2_LIST + KDNB_W2_LIST / 4)
#endif
constexpr int W2_LIST_FLAT = KDNB_W2_LIST_FLAT;  // capacity on planar inputs, where the z array's bytes are free: 120 (-0.3 % against 96, +2 % at 64:
                                                 // profiles/r02_ab_walk_list_capacity.txt)
constexpr int W2_LIST = KDNB_W2_LIST;    // interaction-list capacity (appends come in groups of <= 32); multiple of 4

#ifndef KDNB_W2_UNROLL
#define KDNB_W2_UNROLL 2
#endif
constexpr int W2_UNROLL = KDNB_W2_UNROLL;  // list blocks per drain-loop iteration: 2 is 0.7 % faster than 1 at every size (profiles/r02_ab_walk_unroll.txt)

// 1.875 = the e^2 coefficient of (1 - e)^(-3/2)
__constant__ double W2_C2 = 1.875;

struct __align__(32) Rec32 {
  double a, b, c, d;
};

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// four consecutive list entries: monopoles {cm, m} or leaf particles {p, m} and their lane masks; exact-math contexts
// add 1 = leaf particle per entry.  A multiple of 16 bytes, so that the drain loop addresses everything off one base
// register with 16-byte loads.
template <bool EXACT>
struct W2Blk;
template <>
struct __align__(16) W2Blk<false> {
  double x[4], y[4], m[4];
  uint32_t mask[4];
};
template <>
struct __align__(16) W2Blk<true> {
  double x[4], y[4], m[4];
  uint32_t mask[4];
  uint32_t flag[4];
};
static_assert(sizeof(W2Blk<false>) == 112 && sizeof(W2Blk<true>) == 128, "W2Blk");

template <bool COUNTS>
struct W2CountSmem {
  uint32_t mixmk[32];  // entry masks of the batch's mixed nodes (walk counters only)
};
template <>
struct W2CountSmem<false> {};

// shared memory of one warp: 7296 bytes in the production kernel (no flags, no counter scratch)
template <bool EXACT, bool COUNTS>
struct W2Smem : W2CountSmem<COUNTS> {
  union {
    struct {
      W2Blk<EXACT> blk[W2_LIST / 4];
      double lz[W2_LIST];  // z of the list entries (general inputs only)
    };
    W2Blk<EXACT> blkf[W2_LIST_FLAT / 4];  // planar inputs: no z, the same bytes hold a longer list
  };
  uint32_t snode[W2_STACK + W2_SLACK];
  uint32_t smask[W2_STACK + W2_SLACK];
  union {
    Rec32 mix[32];  // {cx, cy, cz, size^2} of the batch's mixed nodes, compacted
    uint4 lq[32];   // {first slot, num_parts, lane mask, -} of the batch's leaves, compacted
  };
  uint32_t mres[32];  // accept ballots of the batch's mixed nodes
};
static_assert(W2_LIST != 96 || W2_STACK != 304 || sizeof(W2Smem<false, false>) == 7296, "production walk: 28 CTAs x (7296 + 1024 reserved) bytes per SM");

// the reference's formulas (KDNB_FLAG_EXACT_MATH): node -m / (dist_sqr * dist) (array_kd_tree.rs:608), particle
// -m / (dist*dist*dist) (array_particle.rs:72), IEEE sqrt and divide, unfused
__device__ __forceinline__ void interact_exact(double ex, double ey, double ez, double em, bool use, bool is_particle,
                                               double px, double py, double pz, double& ax, double& ay, double& az) {
  if (use) {
    const double dx = __dsub_rn(px, ex), dy = __dsub_rn(py, ey), dz = __dsub_rn(pz, ez);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const double dist = __dsqrt_rn(d2);
    const double den = is_particle ? __dmul_rn(__dmul_rn(dist, dist), dist) : __dmul_rn(d2, dist);
    const double magi = __ddiv_rn(-em, den);
    ax = __dadd_rn(ax, __dmul_rn(magi, dx));
    ay = __dadd_rn(ay, __dmul_rn(magi, dy));
    az = __dadd_rn(az, __dmul_rn(magi, dz));
  }
}

// KDNB_WALK_CPASYNC (compile-time experiment, -DKDNB_WALK_CPASYNC): leaf particles go from global memory straight into
// their list slots with cp.async (LDGSTS, 8 bytes per coordinate: the list is SoA in blocks of four) instead of a
// 32-byte load into registers + four shared-memory stores; the drain waits for the group before it reads the list.
// Measured (profiles/r02_ab_walk_cpasync.txt, parity green): walk 2.081 -> 2.095 ms at N=1M, 20.28 -> 20.45 ms at N=10M —
// the kernel is bound by issue slots, its loads were never exposed, and the asynchronous copies no
// longer overlap the drain that used to run between the load and the stores.  Not the default.
#ifdef KDNB_WALK_CPASYNC
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

// all lanes stream over the list; lane = particle.  Unless this is the group's last drain, only whole blocks of four
// entries are evaluated and the 1-3 entries of an incomplete block move to the front of the list (returned count):
// padding a block with masked-out entries costs a full interaction each, ~2 % of the list slots at 24 drains per group.
#ifdef KDNB_WALK_AB
__device__ int w2_dbg;  // development experiments: 1 = skip the drains, 2 = drain every list twice
#endif
template <bool EXACT, bool COUNTS, bool FLATZ>
__device__ __forceinline__ int drain2(W2Smem<EXACT, COUNTS>& S, int cnt, bool last, int lane, double px, double py, double pz,
                                      double& ax, double& ay, double& az) {
#ifdef KDNB_WALK_CPASYNC
  cp_async_wait_all();
#endif
  __syncwarp();
#ifdef KDNB_WALK_AB
  const int dbg = w2_dbg;
  if (dbg == 1) return 0;
#endif
  if constexpr (EXACT) {
    for (int i = 0; i < cnt; ++i) {
      const W2Blk<EXACT>& B = S.blk[i >> 2];
      const int s = i & 3;
      interact_exact(B.x[s], B.y[s], S.lz[i], B.m[s], (B.mask[s] >> lane) & 1u, B.flag[s] != 0, px, py, pz, ax, ay, az);
    }
  } else {
    // -m d / r^3 without divide or sqrt: y0 = MUFU.RSQ64H estimate (rel. error < 2^-22), e = 1 - d2*y0^2,
    // r^-3 = y0^3 (1 - e)^(-3/2) = y0^3 (1 + 1.5 e + 1.875 e^2 + O(e^3)), O(e^3) < 2^-63.  Four interactions in
    // lock-step, interleaved stage by stage; the list is padded to whole blocks with masked-out entries.  A masked-out
    // lane zeroes the estimate (one 32-bit select on its high word; the low word of an estimate is 0), which makes its
    // contribution exactly -0 * d = no-op and also absorbs d2 == 0 (inf estimate).  (Predicating the accumulating FMAs
    // instead does not pay: ptxas turns a predicated DFMA into DFMA + two FSEL.)
    const uint32_t lanebit = 1u << lane;
#ifdef KDNB_W2_NO_CARRY
    last = true;
#endif
    const int rem = last ? 0 : (cnt & 3);  // entries carried over to the next drain
    cnt -= rem;
    const int nblk = (cnt + 3) >> 2;
    if (lane < 4 * nblk - cnt) {
      const int i = cnt + lane;
      W2Blk<EXACT>& B = (FLATZ ? S.blkf : S.blk)[i >> 2];
      B.x[i & 3] = B.y[i & 3] = B.m[i & 3] = 0.0;
      B.mask[i & 3] = 0u;
      if (!FLATZ) S.lz[i] = 0.0;
    }
    __syncwarp();
    const double* lz = S.lz;
    const W2Blk<EXACT>* B0 = FLATZ ? S.blkf : S.blk;
#ifdef KDNB_WALK_AB
    for (int rep = 0; rep < (dbg == 2 ? 2 : 1); ++rep, lz = S.lz)
#endif
#pragma unroll W2_UNROLL
    for (const W2Blk<EXACT>*B = B0, *E = B0 + nblk; B != E; ++B, lz += 4) {
      double dx[4], dy[4], dz[4], d2[4], y[4], y2[4], ee[4], mq[4], q[4];
      const uint4 m4 = *reinterpret_cast<const uint4*>(B->mask);
      const uint32_t use[4] = {m4.x & lanebit, m4.y & lanebit, m4.z & lanebit, m4.w & lanebit};
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const double2 ex = *reinterpret_cast<const double2*>(&B->x[j]);
        const double2 ey = *reinterpret_cast<const double2*>(&B->y[j]);
        const double2 em = *reinterpret_cast<const double2*>(&B->m[j]);
        dx[j] = __dsub_rn(px, ex.x), dx[j + 1] = __dsub_rn(px, ex.y);
        dy[j] = __dsub_rn(py, ey.x), dy[j + 1] = __dsub_rn(py, ey.y);
        if (!FLATZ) {
          const double2 ez = *reinterpret_cast<const double2*>(&lz[j]);
          dz[j] = __dsub_rn(pz, ez.x), dz[j + 1] = __dsub_rn(pz, ez.y);
        }
        mq[j] = -em.x, mq[j + 1] = -em.y;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) d2[j] = __dmul_rn(dx[j], dx[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) d2[j] = fma(dy[j], dy[j], d2[j]);
      if (!FLATZ) {
#pragma unroll
        for (int j = 0; j < 4; ++j) d2[j] = fma(dz[j], dz[j], d2[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double r = rsqrt_estimate(d2[j]);
        y[j] = __hiloint2double(use[j] ? __double2hiint(r) : 0, 0);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) y2[j] = __dmul_rn(y[j], y[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ee[j] = fma(-d2[j], y2[j], 1.0);
        y[j] = __dmul_rn(y[j], y2[j]);  // y0^3
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        q[j] = fma(W2_C2, ee[j], 1.5);
        mq[j] = __dmul_rn(mq[j], y[j]);  // -m * y0^3
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) ee[j] = __dmul_rn(mq[j], ee[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) mq[j] = fma(ee[j], q[j], mq[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ax = fma(mq[j], dx[j], ax);
        ay = fma(mq[j], dy[j], ay);
        if (!FLATZ) az = fma(mq[j], dz[j], az);
      }
    }
    if (rem) {  // (cnt is a multiple of 4 here, and > 0: a drain that is not the last one is triggered by a full list)
      __syncwarp();
      if (lane < rem) {
        W2Blk<EXACT>* Bw = FLATZ ? S.blkf : S.blk;
        const W2Blk<EXACT>& Bs = Bw[cnt >> 2];
        Bw[0].x[lane] = Bs.x[lane], Bw[0].y[lane] = Bs.y[lane], Bw[0].m[lane] = Bs.m[lane];
        Bw[0].mask[lane] = Bs.mask[lane];
        if (!FLATZ) S.lz[lane] = S.lz[cnt + lane];
      }
    }
    __syncwarp();
    return rem;
  }
  __syncwarp();
  return 0;
}

// store list entry i
template <bool EXACT, bool COUNTS, bool FLATZ>
__device__ __forceinline__ void list_put(W2Smem<EXACT, COUNTS>& S, int i, double x, double y, double z, double m, uint32_t mask,
                                         uint32_t flag) {
  W2Blk<EXACT>& B = (FLATZ ? S.blkf : S.blk)[i >> 2];
  const int s = i & 3;
  B.x[s] = x, B.y[s] = y, B.m[s] = m;
  B.mask[s] = mask;
  if (!FLATZ) S.lz[i] = z;
  if constexpr (EXACT) B.flag[s] = flag;
}

enum : int { W2_NONE = 0, W2_FAR = 1, W2_NEAR = 2, W2_MIXED = 3, W2_LEAF = 4 };

template <bool EXACT, bool COUNTS, bool PEER, bool FLATZ>
__device__ __forceinline__ void walk2_body(W2Smem<EXACT, COUNTS>& S, const WNode* __restrict__ nodes,
                                           const PosM* __restrict__ posm, double* __restrict__ acc_t,
                                           uint32_t slot_begin, uint32_t slot_end, double theta2,
                                           unsigned long long* __restrict__ wcounts, const P2P& p2p, int lshift,
                                           const uint32_t* __restrict__ gorder, uint32_t* __restrict__ gcost,
                                           const uint32_t* __restrict__ seed) {
  const int lane = threadIdx.x;
  const uint32_t lt = (1u << lane) - 1u;
  // CTA -> group of 32 tree slots: in launch order, or heaviest first by the previous step's work (walk.cu)
  const uint32_t group = gorder ? gorder[blockIdx.x] : blockIdx.x;
  const uint32_t base = slot_begin + group * 32u;
  uint32_t work = 0;  // drained list entries + 4 per batch: proportional to the issue slots this group used
  const uint32_t slot = base + lane;
  const bool valid_p = slot < slot_end;
  const bool warp_has_work = base < slot_end;
  if (!PEER && !warp_has_work) return;  // (peer mode: no early exit, every CTA joins the end-of-kernel handshake)
  double px = 0.0, py = 0.0, pz = 0.0, ax = 0.0, ay = 0.0, az = 0.0;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  unsigned long long cv = 0, ca = 0, cl = 0, cp = 0;
  if (valid_p) {
    const PosM me = posm[slot];
    px = me.x, py = me.y, pz = me.z;
    lo[0] = hi[0] = me.x;
    lo[1] = hi[1] = me.y;
    lo[2] = hi[2] = me.z;
  }
  constexpr int ND = FLATZ ? 2 : 3;
  constexpr int LCAP = FLATZ ? W2_LIST_FLAT : W2_LIST;  // interaction-list capacity
  // Box of the group as centre and half extent.  hs is the half extent INFLATED by 2^-46 of the box's scale: the
  // distance of a node to the box along one axis is then bounded from below by |c - mid| - hs and from above by
  // |c - mid| + hs whatever the roundings of mid, hs and the subtractions (each <= 2^-52 of that scale), also where
  // |c - mid| and hs cancel; relative errors elsewhere are covered by the 1e-9 margins of the far / near tests.
  double mid[3] = {0.0, 0.0, 0.0}, hs[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < ND; ++k) {
    lo[k] = warp_min(lo[k]);
    hi[k] = warp_max(hi[k]);
    mid[k] = 0.5 * lo[k] + 0.5 * hi[k];
    const double half = 0.5 * hi[k] - 0.5 * lo[k];
    hs[k] = half + 1.4210854715202004e-14 * (fabs(mid[k]) + half);
  }
  {
    const uint32_t m0 = __ballot_sync(0xffffffffu, valid_p);
    if (lane == 0) {
      S.snode[0] = 0;
      S.smask[0] = m0;
    }
  }
  __syncwarp();
  const double tfar = theta2 * (1.0 - 1e-9), tnear = theta2 * (1.0 + 1e-9);
  const int kmask = (1 << lshift) - 1, lpr = 32 >> lshift;

  int ln = 0;
  int sp = warp_has_work ? 1 : 0;
  // ---- the common descent.  Every group would start with five nearly empty batches (1, 2, 4, 8, 16 nodes) that open
  // the same top of the tree.  seed[] (plan(): node indices of depths 0-5 in heap order, depths 0-4 all internal):
  // lanes 0-30 test the 31 nodes of depths 0-4 against the group's box; if every particle opens all of them (the rule
  // with theta < 1), the traversal starts from the 32 nodes of depth 5.  Otherwise it starts at the root as before.
  if (seed != nullptr && warp_has_work) {
    const uint32_t nd = seed[lane];  // (lane 31 reads the first node of depth 5 and is not part of the vote)
    const Rec32* rec = reinterpret_cast<const Rec32*>(nodes + nd);
    const int4 info = __ldg(reinterpret_cast<const int4*>(rec + 1));
    const Rec32 c = rec[0];
    const double size2 = __hiloint2double(info.y, info.x);
    const double cc[3] = {c.a, c.b, c.c};
    double dmax2 = 0.0;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      const double df = fabs(cc[k] - mid[k]) + hs[k];
      dmax2 = fma(df, df, dmax2);
    }
    const bool opened = lane == 31 || size2 >= tnear * dmax2;
    if (__all_sync(0xffffffffu, opened)) {
      S.snode[lane] = seed[31 + lane];
      S.smask[lane] = __ballot_sync(0xffffffffu, valid_p);
      sp = 32;
      if (COUNTS && valid_p) cv += 31;  // every particle tested (and opened) the 31 nodes above
      __syncwarp();
    }
  }
  while (sp > 0) {
    // ---- pop a batch: lane l takes entry sp+l after the pop (order inside a batch is irrelevant); lanes beyond the
    // batch re-read its first entry (no divergence, no dead values) and drop out after the classification
    const int room = W2_STACK - sp;
    const int nb = min(min(sp, 32), max(1, room));
    sp -= nb;
    work += 4u;
    const bool has = lane < nb;
    const int at = sp + (has ? lane : 0);
    const uint32_t node = S.snode[at];
    const uint32_t mk = S.smask[at];
    const Rec32* rec = reinterpret_cast<const Rec32*>(nodes + node);
    const int4 info = __ldg(reinterpret_cast<const int4*>(rec + 1));  // size2, (a, b)
    const Rec32 c = rec[0];  // cx, cy, cz, m (meaningless for a leaf; loaded alongside so the two sectors travel together)
    const uint32_t na = (uint32_t)info.z, nbits = (uint32_t)info.w;
    const double size2 = __hiloint2double(info.y, info.x);
    int kind;
    {
      double dmin2 = 0.0, dmax2 = 0.0;
      const double cc[3] = {c.a, c.b, c.c};
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        const double t = fabs(cc[k] - mid[k]);
        double dn = t - hs[k];
        // max(dn, 0) on the sign: a negative dn becomes a denormal (high word 0), whose square is 0
        dn = __hiloint2double(max(__double2hiint(dn), 0), __double2loint(dn));
        const double df = t + hs[k];
        dmin2 = fma(dn, dn, dmin2);
        dmax2 = fma(df, df, dmax2);
      }
      kind = W2_MIXED;
      if (size2 < tfar * dmin2) kind = W2_FAR;
      else if (size2 >= tnear * dmax2) kind = W2_NEAR;
      if (!(nbits & WN_INTERNAL)) kind = W2_LEAF;
      if (!has) kind = W2_NONE;
    }
    __syncwarp();
    if (COUNTS) {  // every particle in an entry's mask tests that node (and accepts it when it is far)
      for (int s = 0; s < nb; ++s) {
        const int ks = __shfl_sync(0xffffffffu, kind, s);
        const uint32_t ms = __shfl_sync(0xffffffffu, mk, s);
        const unsigned long long bit = (ms >> lane) & 1u;
        if (ks == W2_FAR || ks == W2_NEAR) cv += bit;
        if (ks == W2_FAR) ca += bit;
        if (ks == W2_LEAF) cl += bit;
      }
    }
    // ---- mixed nodes: the reference's test, one particle per lane (array_kd_tree.rs:601-606)
    uint32_t amask = kind == W2_FAR ? mk : 0u;   // lanes that accept this lane's node
    uint32_t omask = kind == W2_NEAR ? mk : 0u;  // lanes that open it
    const uint32_t bal_mixed = __ballot_sync(0xffffffffu, kind == W2_MIXED);
    if (bal_mixed) {
      const int mrank = __popc(bal_mixed & lt);
      if (kind == W2_MIXED) {
        Rec32 t;
        t.a = c.a, t.b = c.b, t.c = c.c, t.d = size2;
        S.mix[mrank] = t;
        if constexpr (COUNTS) S.mixmk[mrank] = mk;
      }
      __syncwarp();
      const int nm = __popc(bal_mixed);
      for (int k = 0; k < nm; ++k) {
        const Rec32 q = S.mix[k];
        const double dx = __dsub_rn(px, q.a), dy = __dsub_rn(py, q.b);
        double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));  // :604, left to right
        if (!FLATZ) {
          const double dz = __dsub_rn(pz, q.c);
          d2 = __dadd_rn(d2, __dmul_rn(dz, dz));
        }
        const bool accept = q.d < __dmul_rn(theta2, d2);  // :606
        const uint32_t b = __ballot_sync(0xffffffffu, accept);
        if (lane == 0) S.mres[k] = b;
        if constexpr (COUNTS) {
          const bool in = (S.mixmk[k] >> lane) & 1u;
          cv += in;
          ca += in && accept;
        }
      }
      __syncwarp();
      if (kind == W2_MIXED) {
        const uint32_t am = S.mres[mrank];
        amask = am & mk;
        omask = mk & ~am;
      }
      __syncwarp();  // S.mix is reused as the leaf queue below
    }
    // ---- far + accepted: append the monopole with the mask of the accepting lanes
    {
      const bool mine = amask != 0;
      const uint32_t bal = __ballot_sync(0xffffffffu, mine);
      if (bal) {
        const int add = __popc(bal);
        if (ln + add > LCAP) {
          const int left = drain2<EXACT, COUNTS, FLATZ>(S, ln, false, lane, px, py, pz, ax, ay, az);
          work += (uint32_t)(ln - left);
          ln = left;
        }
        if (mine) list_put<EXACT, COUNTS, FLATZ>(S, ln + __popc(bal & lt), c.a, c.b, c.c, c.d, amask, 0u);
        ln += add;
      }
    }
    // ---- near + still open: push both children with the mask of the opening lanes
    {
      const bool mine = omask != 0;
      const uint32_t bal = __ballot_sync(0xffffffffu, mine);
      if (mine) {
        const int i = sp + 2 * __popc(bal & lt);
        S.snode[i] = na;            // right
        S.snode[i + 1] = node + 1;  // left (the next record)
        S.smask[i] = omask;
        S.smask[i + 1] = omask;
      }
      sp += 2 * __popc(bal);
    }
    // ---- leaves: lane -> (leaf, k), 32 >> lshift leaves per round
    const uint32_t bal_leaf = __ballot_sync(0xffffffffu, kind == W2_LEAF);
    if (bal_leaf) {
      const int nl = __popc(bal_leaf);
      if (kind == W2_LEAF) S.lq[__popc(bal_leaf & lt)] = make_uint4(na, nbits, mk, 0u);
      __syncwarp();
      for (int r = 0; r < nl; r += lpr) {
        const int li = r + (lane >> lshift);
        const uint32_t k = (uint32_t)(lane & kmask);
        uint4 L = make_uint4(0u, 0u, 0u, 0u);
        if (li < nl) L = S.lq[li];
        const uint32_t j = L.x + k;
        const uint32_t t = j - base;  // the lane that owns particle j, if it is one of ours
        const uint32_t m = t < 32u ? (L.z & ~(1u << t)) : L.z;
        const bool valid = k < L.y && m != 0u;
        const uint32_t bal = __ballot_sync(0xffffffffu, valid);
        const int add = __popc(bal);
#ifdef KDNB_WALK_CPASYNC
        if (ln + add > LCAP) {
          const int left = drain2<EXACT, COUNTS, FLATZ>(S, ln, false, lane, px, py, pz, ax, ay, az);
          work += (uint32_t)(ln - left);
          ln = left;
        }
        if (valid) {
          const int i = ln + __popc(bal & lt);
          W2Blk<EXACT>& B = (FLATZ ? S.blkf : S.blk)[i >> 2];
          const int s = i & 3;
          const PosM* g = posm + j;
          cp_async8(&B.x[s], &g->x);
          cp_async8(&B.y[s], &g->y);
          cp_async8(&B.m[s], &g->m);
          if (!FLATZ) cp_async8(&S.lz[i], &g->z);
          B.mask[s] = m;
          if constexpr (EXACT) B.flag[s] = 1u;
        }
        ln += add;
#else
        PosM qv;
        if (valid) qv = posm[j];
        if (ln + add > LCAP) {
          const int left = drain2<EXACT, COUNTS, FLATZ>(S, ln, false, lane, px, py, pz, ax, ay, az);
          work += (uint32_t)(ln - left);
          ln = left;
        }
        if (valid) list_put<EXACT, COUNTS, FLATZ>(S, ln + __popc(bal & lt), qv.x, qv.y, qv.z, qv.m, m, 1u);
        ln += add;
#endif
        if (COUNTS) {
          for (int s = 0; s < 32; ++s) {
            const uint32_t ms = __shfl_sync(0xffffffffu, valid ? m : 0u, s);
            cp += (ms >> lane) & 1u;
          }
        }
      }
    }
    __syncwarp();
  }
  drain2<EXACT, COUNTS, FLATZ>(S, ln, true, lane, px, py, pz, ax, ay, az);
  work += (uint32_t)ln;
  if (gcost && lane == 0 && warp_has_work) gcost[group] = work;

  // ---- results.  Single GPU / NCCL mode: tree-ordered accelerations into the local acc_t.  Peer mode: the same
  // 24 bytes go straight into EVERY rank's acc_t over NVLink (this rank's shard of everyone's copy), followed by a
  // system-scope fence; the last CTA to finish then raises this rank's flag on every peer (p2p_wait_kernel consumes it).
  const bool peer = PEER && p2p.world > 1;
  uint32_t epoch = 0;
  uint64_t boff = 0;
  if (peer) {
    epoch = *p2p.epoch;
    boff = (uint64_t)(epoch & 1u) * p2p.stride;
  }
  if (valid_p) {
    if (!peer) {
      acc_t[3ull * slot + 0] = ax;
      acc_t[3ull * slot + 1] = ay;
      acc_t[3ull * slot + 2] = az;
    } else {
      for (int r = 0; r < p2p.world; ++r) {
        double* dst = p2p.acc[r] + boff + 3ull * slot;
        dst[0] = ax;
        dst[1] = ay;
        dst[2] = az;
      }
    }
    if (COUNTS) {
      wcounts[4ull * slot + 0] = cv;
      wcounts[4ull * slot + 1] = ca;
      wcounts[4ull * slot + 2] = cl;
      wcounts[4ull * slot + 3] = cp;
    }
  }
  if (PEER && peer) {
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
      const uint32_t done = atomicAdd(p2p.cta_done, 1u);
      if (done == gridDim.x - 1) {
        *p2p.cta_done = 0;
        __threadfence_system();
        for (int r = 0; r < p2p.world; ++r) *reinterpret_cast<volatile uint32_t*>(p2p.flags[r] + p2p.rank) = epoch + 1u;
      }
    }
  }
}

// grid = ceil((slot_end - slot_begin) / 32) CTAs of one warp
template <bool EXACT, bool COUNTS, bool PEER, int MINB>
__global__ void __launch_bounds__(32, MINB)
walk2_kernel(const WNode* __restrict__ nodes, const PosM* __restrict__ posm, double* __restrict__ acc_t,
             uint32_t slot_begin, uint32_t slot_end, double theta2, unsigned long long* __restrict__ wcounts,
             P2P p2p, const uint32_t* __restrict__ flat, int lshift, const uint32_t* __restrict__ gorder,
             uint32_t* __restrict__ gcost, const uint32_t* __restrict__ seed) {
  pdl_sync();
  __shared__ W2Smem<EXACT, COUNTS> S;
  if (!EXACT && !COUNTS && flat[3])
    walk2_body<EXACT, COUNTS, PEER, true>(S, nodes, posm, acc_t, slot_begin, slot_end, theta2, wcounts, p2p, lshift, gorder, gcost, seed);
  else
    walk2_body<EXACT, COUNTS, PEER, false>(S, nodes, posm, acc_t, slot_begin, slot_end, theta2, wcounts, p2p, lshift, gorder, gcost, seed);
}

// Heaviest-first launch order for the next walk (one CTA): gorder = the groups sorted by descending gcost, by a
// counting sort on 1024 cost classes.  Every walk launch takes ~0.1-0.2 ms beyond its issue-slot work whatever the grid
// (profiles/README.md): it ends with whatever its last CTAs happen to be, the groups differ by up to 2x in work (std
// 18 % of the mean), and the hardware hands CTAs to the SMs in index order.  Dealing the heavy groups first shortens
// that drain of the machine: measured 2.187 -> 2.107 ms at 31251 CTAs (N = 1M) and 0.370 -> 0.303 ms at the 3907 CTAs
// of a 1/8 shard (profiles/r01_ab_walk_lpt.txt; tests/devtools/launch_model.py predicted -6 % and -15 %).  The order
// inside a class is whatever the atomics give: results do not depend on which CTA computes which group.
__global__ void __launch_bounds__(1024) walk_order_kernel(const uint32_t* __restrict__ gcost, uint32_t* __restrict__ gorder,
                                                          uint32_t ngroups) {
  pdl_sync();
  __shared__ uint32_t cls[1024];
  __shared__ uint32_t red[32];
  const uint32_t t = threadIdx.x;
  uint32_t mx = 1u;
  for (uint32_t g = t; g < ngroups; g += 1024u) mx = max(mx, gcost[g]);
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((t & 31u) == 0u) red[t >> 5] = mx;
  cls[t] = 0u;
  __syncthreads();
  mx = red[t & 31u];
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const double scale = 1023.0 / (double)mx;
  // class 0 = heaviest
  for (uint32_t g = t; g < ngroups; g += 1024u) atomicAdd(&cls[1023u - (uint32_t)((double)gcost[g] * scale)], 1u);
  __syncthreads();
  // exclusive scan of the 1024 class counts (one per thread)
  const uint32_t mine = cls[t];
  uint32_t incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)(t & 31u) >= o) incl += v;
  }
  __syncthreads();
  if ((t & 31u) == 31u) red[t >> 5] = incl;
  __syncthreads();
  uint32_t before = 0u;
  for (uint32_t w = 0; w < (t >> 5); ++w) before += red[w];
  __syncthreads();
  cls[t] = before + incl - mine;
  __syncthreads();
  for (uint32_t g = t; g < ngroups; g += 1024u) {
    const uint32_t pos = atomicAdd(&cls[1023u - (uint32_t)((double)gcost[g] * scale)], 1u);
    gorder[pos] = g;
  }
}

}  // namespace kdnb
