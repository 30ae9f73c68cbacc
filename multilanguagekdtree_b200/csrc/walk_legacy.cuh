// walk_legacy.cuh — the FIRST walk kernel of round 1 (mixed nodes and leaves handled one at a time by the whole warp),
// kept selectable for A/B timing (KDNB_WALK_CFG=1), plus the helpers walk2.cuh shares with it (Rec32, interact<>,
// warp_min / warp_max).  The kernel that runs by default is walk2.cuh's.
//
// theta-criterion force walk, calc_accel / accel_recur of the reference
// (Parallel/RustVersion/src/array_kd_tree.rs:585-621) with calc_pp_accel (array_particle.rs:67-76) in the leaves.
//
// One warp owns 32*PPL consecutive TREE-ORDERED particles (a compact patch of ~10 leaves), PPL per lane, and keeps a
// shared-memory stack of frontier entries (node, one lane mask per 32 particles).  Per round it pops up to 32
// entries and classifies them ONE NODE PER LANE against the bounding box of its particles:
//   far   : size^2 <  theta^2 * dmin^2 * (1 - 1e-9)  -> every particle in the entry's masks accepts the node
//   near  : size^2 >= theta^2 * dmax^2 * (1 + 1e-9)  -> every particle opens it: both children are pushed
//   mixed : otherwise, and leaves                    -> handled one node at a time by the whole warp
// dmin/dmax are the distances from the node's centre of mass to the box; the 1e-9 margin dwarfs the <= 1e-15
// rounding of either side, so "far"/"near" provably agree with the reference's per-particle test
//       size*size < (THETA*THETA) * dist_sqr            (array_kd_tree.rs:606)
// For a mixed node every lane evaluates exactly that test for its own particles, with the reference's unfused
// operation order, and __ballot_sync splits the masks into accepted and still-open lanes.  Every particle thus
// accepts / opens precisely the nodes the reference's recursion does (checked by KDNB_FLAG_WALK_COUNTS against the
// oracle: per-particle counts of tests, accepts, leaf visits and pair interactions are identical).
//
// Forces are not evaluated during the traversal: accepted monopoles {cm, m} and leaf particles (minus the lane that
// owns the particle: leaf_parts[i] != p, :590) are appended with their lane mask to a per-warp, per-32-particle
// interaction list in shared memory, which is drained by a branch-free, 4-way unrolled loop: 16 FP64 instructions
// per interaction, broadcast shared-memory loads, no global loads.  About 3/4 of the nodes a warp touches are
// far / near (profiles/README.md), so the serial per-node work — which dominated the first versions of this kernel —
// shrinks to the mixed nodes, and the drain loop runs on lists that skip 32-particle halves nobody in them needs.
//
// Accumulation is a running f64 sum per particle (the reference combines pairwise along the recursion, :611-613);
// the difference is summation order only and is covered by the stated 1e-12 tolerance.
#pragma once
#include <cstdlib>

#include "ctx.cuh"

namespace kdnb {

constexpr int WALK_STACK = 320;  // soft capacity: batches shrink as the stack fills (see nb below)
constexpr int WALK_SLACK = 32;   // depth-first tail when the stack is at capacity (tree depth <= 27 at 1e8 particles)
constexpr int WALK_LIST = 64;    // interaction-list capacity per 32 particles (>= 32 + largest MAX_PARTS)

struct __align__(32) Rec32 {
  double a, b, c, d;
};

// -m / r^3 without divide or sqrt: y0 = MUFU.RSQ64H estimate (rel. error < 2^-22), e = 1 - d2*y0^2,
// r^-3 = y0^3 * (1 - e)^(-3/2) = y0^3 * (1 + 1.5 e + 1.875 e^2 + O(e^3)); O(e^3) < 2^-63.  No special cases:
// callers discard the result by select when the pair is masked out (d2 == 0 gives NaN there).
__device__ __forceinline__ double neg_m_over_r3_fast(double mneg, double d2) {
  const double y = rsqrt_estimate(d2);
  const double y2 = __dmul_rn(y, y);
  const double e = fma(-d2, y2, 1.0);
  const double y3 = __dmul_rn(y, y2);
  const double q = fma(1.875, e, 1.5);
  const double mq = __dmul_rn(mneg, y3);
  return fma(__dmul_rn(mq, e), q, mq);
}

template <int PPL, int WALK_WARPS>
struct WalkSmem {
  uint32_t snode[WALK_WARPS][WALK_STACK + WALK_SLACK];
  uint32_t smask[WALK_WARPS][PPL][WALK_STACK + WALK_SLACK];
  Rec32 lpos[WALK_WARPS][PPL][WALK_LIST];   // {x, y, z, m} of a monopole or of a leaf particle
  uint2 lmask[WALK_WARPS][PPL][WALK_LIST];  // {lane mask, is_particle}
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

// all lanes stream over one 32-particle list
template <bool EXACT, bool COUNTS, int DW>
__device__ __forceinline__ void drain_list(const Rec32* __restrict__ lpos, const uint2* __restrict__ lmask, int cnt,
                                           int lane, double px, double py, double pz, double& ax, double& ay,
                                           double& az, unsigned long long& cp) {
  __syncwarp();
  if (EXACT || COUNTS) {
    for (int i = 0; i < cnt; ++i) {
      const Rec32 e = lpos[i];
      const uint2 mk = lmask[i];
      const bool use = (mk.x >> lane) & 1u;
      interact<EXACT>(e, use, mk.y != 0, px, py, pz, ax, ay, az);
      if (COUNTS) cp += (use && mk.y) ? 1 : 0;
    }
  } else {
    // DW interactions in lock-step: the FP64 chain of one interaction is ~11 instructions deep, so the independent
    // chains are interleaved by hand, stage by stage (left to the compiler they were emitted one after another and
    // the FP64 pipe idled on its own latency: profiles/README.md).  The list is padded to a multiple of DW with
    // masked-out entries.
    const uint32_t lanebit = 1u << lane;
    const int padded = (cnt + DW - 1) / DW * DW;
    if (lane < padded - cnt) {
      Rec32 z;
      z.a = z.b = z.c = z.d = 0.0;
      const_cast<Rec32*>(lpos)[cnt + lane] = z;
      const_cast<uint2*>(lmask)[cnt + lane] = make_uint2(0u, 0u);
    }
    __syncwarp();
    for (int i = 0; i < padded; i += DW) {
      double dx[DW], dy[DW], dz[DW], d2[DW], y[DW], y2[DW], ee[DW], mq[DW], q[DW];
      uint32_t use[DW];
      {  // the DW masks in two 16-byte loads ({mask, flag} pairs)
#pragma unroll
        for (int j = 0; j < DW; j += 2) {
          const uint4 m = *reinterpret_cast<const uint4*>(&lmask[i + j]);
          use[j] = m.x & lanebit;
          use[j + 1] = m.z & lanebit;
        }
      }
#pragma unroll
      for (int j = 0; j < DW; ++j) {
        const Rec32 e = lpos[i + j];
        dx[j] = __dsub_rn(px, e.a);
        dy[j] = __dsub_rn(py, e.b);
        dz[j] = __dsub_rn(pz, e.c);
        mq[j] = -e.d;
      }
#pragma unroll
      for (int j = 0; j < DW; ++j) d2[j] = __dmul_rn(dx[j], dx[j]);
#pragma unroll
      for (int j = 0; j < DW; ++j) d2[j] = fma(dy[j], dy[j], d2[j]);
#pragma unroll
      for (int j = 0; j < DW; ++j) d2[j] = fma(dz[j], dz[j], d2[j]);
#pragma unroll
      for (int j = 0; j < DW; ++j) y[j] = rsqrt_estimate(d2[j]);
#pragma unroll
      for (int j = 0; j < DW; ++j) y2[j] = __dmul_rn(y[j], y[j]);
#pragma unroll
      for (int j = 0; j < DW; ++j) {
        ee[j] = fma(-d2[j], y2[j], 1.0);
        y[j] = __dmul_rn(y[j], y2[j]);  // y0^3
      }
#pragma unroll
      for (int j = 0; j < DW; ++j) {
        q[j] = fma(1.875, ee[j], 1.5);
        mq[j] = __dmul_rn(mq[j], y[j]);  // -m * y0^3
      }
#pragma unroll
      for (int j = 0; j < DW; ++j) ee[j] = __dmul_rn(mq[j], ee[j]);
#pragma unroll
      for (int j = 0; j < DW; ++j) {
        const double magi = fma(ee[j], q[j], mq[j]);
        mq[j] = use[j] ? magi : 0.0;  // (ptxas turns predicated DFMAs into DFMA + 2 FSEL each: select once here)
      }
#pragma unroll
      for (int j = 0; j < DW; ++j) {
        ax = fma(mq[j], dx[j], ax);
        ay = fma(mq[j], dy[j], ay);
        az = fma(mq[j], dz[j], az);
      }
    }
  }
  __syncwarp();
}

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

enum : int { K_NONE = 0, K_FAR = 1, K_NEAR = 2, K_SERIAL = 3 };

template <int PPL, int WALK_THREADS, int MINB, bool EXACT, bool COUNTS, bool PEER, int DW = 4>
__global__ void __launch_bounds__(WALK_THREADS, MINB)
walk_kernel(const WNode* __restrict__ nodes, const PosM* __restrict__ posm, double* __restrict__ acc_t,
            uint32_t slot_begin, uint32_t slot_end, double theta2, unsigned long long* __restrict__ wcounts,
            P2P p2p) {
  pdl_sync();
  constexpr int WALK_WARPS = WALK_THREADS / 32;
  __shared__ WalkSmem<PPL, WALK_WARPS> S;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t base = slot_begin + (blockIdx.x * WALK_WARPS + w) * (32 * PPL);
  uint32_t slot[PPL], mk[PPL];
  int ln[PPL];
  double px[PPL], py[PPL], pz[PPL], ax[PPL], ay[PPL], az[PPL];
  unsigned long long cv[PPL], ca[PPL], cl[PPL], cp[PPL];
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  uint32_t* snode = S.snode[w];
#pragma unroll
  for (int u = 0; u < PPL; ++u) {
    slot[u] = base + u * 32 + lane;
    const bool valid = slot[u] < slot_end;
    const uint32_t m0 = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) S.smask[w][u][0] = m0;
    px[u] = py[u] = pz[u] = 0.0;
    if (valid) {
      const PosM me = posm[slot[u]];
      px[u] = me.x;
      py[u] = me.y;
      pz[u] = me.z;
      lo[0] = fmin(lo[0], me.x), hi[0] = fmax(hi[0], me.x);
      lo[1] = fmin(lo[1], me.y), hi[1] = fmax(hi[1], me.y);
      lo[2] = fmin(lo[2], me.z), hi[2] = fmax(hi[2], me.z);
    }
    ax[u] = ay[u] = az[u] = 0.0;
    cv[u] = ca[u] = cl[u] = cp[u] = 0;
    ln[u] = 0;
  }
  const bool warp_has_work = base < slot_end;
  if (!PEER && !warp_has_work) return;  // (peer mode: no early exit, every warp joins the end-of-kernel handshake)
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    lo[k] = warp_min(lo[k]);
    hi[k] = warp_max(hi[k]);
  }
  if (lane == 0) snode[0] = 0;
  __syncwarp();
  const double far_margin = 1.0 - 1e-9, near_margin = 1.0 + 1e-9;

  int sp = warp_has_work ? 1 : 0;
  while (sp > 0) {
    // ---- pop a batch: lane l takes entry sp+l after the pop (order inside a batch is irrelevant)
    const int room = WALK_STACK - sp;
    const int nb = min(min(sp, 32), max(1, room));
    sp -= nb;
    const bool has = lane < nb;
    uint32_t node = 0, na = 0, nbits = 0;
    int kind = K_NONE;
    Rec32 c;
    c.a = c.b = c.c = c.d = 0.0;
#pragma unroll
    for (int u = 0; u < PPL; ++u) mk[u] = 0;
    if (has) {
      node = snode[sp + lane];
#pragma unroll
      for (int u = 0; u < PPL; ++u) mk[u] = S.smask[w][u][sp + lane];
      const Rec32* rec = reinterpret_cast<const Rec32*>(nodes + node);
      const int4 info = __ldg(reinterpret_cast<const int4*>(rec + 1));  // size2, (a, b)
      na = (uint32_t)info.z;
      nbits = (uint32_t)info.w;
      kind = K_SERIAL;
      if (nbits & WN_INTERNAL) {
        c = rec[0];  // cx, cy, cz, m
        const double size2 = __hiloint2double(info.y, info.x);
        double dmin2 = 0.0, dmax2 = 0.0;
        const double cc[3] = {c.a, c.b, c.c};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double below = lo[k] - cc[k], above = cc[k] - hi[k];  // > 0 when the centre is outside the box
          const double dn = fmax(0.0, fmax(below, above));
          const double df = fmax(fabs(below), fabs(above));  // = max(|c-lo|, |c-hi|)
          dmin2 = fma(dn, dn, dmin2);
          dmax2 = fma(df, df, dmax2);
        }
        if (size2 < theta2 * dmin2 * far_margin) kind = K_FAR;
        else if (size2 >= theta2 * dmax2 * near_margin) kind = K_NEAR;
      }
    }
    __syncwarp();
    if (COUNTS) {  // every particle in an entry's mask tests that node (and accepts it when it is far)
      for (int s = 0; s < nb; ++s) {
        const int ks = __shfl_sync(0xffffffffu, kind, s);
        const uint32_t nbs = __shfl_sync(0xffffffffu, nbits, s);
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
          const uint32_t ms = __shfl_sync(0xffffffffu, mk[u], s);
          const unsigned long long bit = (ms >> lane) & 1u;
          if (ks == K_FAR || ks == K_NEAR) cv[u] += bit;
          if (ks == K_FAR) ca[u] += bit;
          if (ks == K_SERIAL && !(nbs & WN_INTERNAL)) cl[u] += bit;
        }
      }
    }
    // ---- far nodes: append the monopole to the lists of the 32-particle halves that hold accepting particles
    if (__any_sync(0xffffffffu, kind == K_FAR)) {
#pragma unroll
      for (int u = 0; u < PPL; ++u) {
        const bool mine = kind == K_FAR && mk[u] != 0;
        const uint32_t bal = __ballot_sync(0xffffffffu, mine);
        const int add = __popc(bal);
        if (ln[u] + add > WALK_LIST) {
          drain_list<EXACT, COUNTS, DW>(S.lpos[w][u], S.lmask[w][u], ln[u], lane, px[u], py[u], pz[u], ax[u], ay[u], az[u], cp[u]);
          ln[u] = 0;
        }
        if (mine) {
          const int i = ln[u] + __popc(bal & lt);
          S.lpos[w][u][i] = c;
          S.lmask[w][u][i] = make_uint2(mk[u], 0u);
        }
        ln[u] += add;
      }
    }
    // ---- near nodes: push both children with the same masks
    {
      const uint32_t bal = __ballot_sync(0xffffffffu, kind == K_NEAR);
      if (kind == K_NEAR) {
        const int i = sp + 2 * __popc(bal & lt);
        snode[i] = na;            // right
        snode[i + 1] = node + 1;  // left (the next record)
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
          S.smask[w][u][i] = mk[u];
          S.smask[w][u][i + 1] = mk[u];
        }
      }
      sp += 2 * __popc(bal);
    }
    // ---- leaves and mixed nodes: one at a time, all lanes
    uint32_t ser = __ballot_sync(0xffffffffu, kind == K_SERIAL);
    while (ser) {
      const int src = __ffs(ser) - 1;
      ser &= ser - 1;
      const uint32_t nd = __shfl_sync(0xffffffffu, node, src);
      const uint32_t a_s = __shfl_sync(0xffffffffu, na, src);
      const uint32_t b_s = __shfl_sync(0xffffffffu, nbits, src);
      uint32_t ms[PPL];
#pragma unroll
      for (int u = 0; u < PPL; ++u) ms[u] = __shfl_sync(0xffffffffu, mk[u], src);
      if (!(b_s & WN_INTERNAL)) {
        // leaf: its particles go to the lists with the masks of the lanes that reached it, minus the owner (:590)
        const int cnt = (int)b_s;
        Rec32 r;
        r.a = r.b = r.c = r.d = 0.0;
        const uint32_t j = a_s + lane;
        if (lane < cnt) {
          const PosM q = posm[j];
          r.a = q.x, r.b = q.y, r.c = q.z, r.d = q.m;
        }
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
          if (ms[u] == 0) continue;
          if (ln[u] + cnt > WALK_LIST) {
            drain_list<EXACT, COUNTS, DW>(S.lpos[w][u], S.lmask[w][u], ln[u], lane, px[u], py[u], pz[u], ax[u], ay[u], az[u], cp[u]);
            ln[u] = 0;
          }
          if (lane < cnt) {
            const uint32_t t = j - (base + u * 32);  // the lane that owns particle j, if it is one of ours
            const uint32_t m = t < 32u ? (ms[u] & ~(1u << t)) : ms[u];
            S.lpos[w][u][ln[u] + lane] = r;
            S.lmask[w][u][ln[u] + lane] = make_uint2(m, 1u);
          }
          ln[u] += cnt;
        }
      } else {
        // mixed node: the reference's test, per particle (array_kd_tree.rs:601-606)
        const Rec32* rec = reinterpret_cast<const Rec32*>(nodes + nd);
        const Rec32 cs = rec[0];
        const double size2 = __ldg(reinterpret_cast<const double*>(rec + 1));
        uint32_t open_any = 0;
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
          const bool in = (ms[u] >> lane) & 1u;
          const double dx = __dsub_rn(px[u], cs.a), dy = __dsub_rn(py[u], cs.b), dz = __dsub_rn(pz[u], cs.c);
          const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));  // :604
          const bool accept = in && (size2 < __dmul_rn(theta2, d2));                                        // :606
          const uint32_t am = __ballot_sync(0xffffffffu, accept);
          if (COUNTS) {
            cv[u] += in;
            ca[u] += accept;
          }
          ms[u] &= ~am;
          open_any |= ms[u];
          if (am) {
            if (ln[u] + 1 > WALK_LIST) {
              drain_list<EXACT, COUNTS, DW>(S.lpos[w][u], S.lmask[w][u], ln[u], lane, px[u], py[u], pz[u], ax[u], ay[u], az[u], cp[u]);
              ln[u] = 0;
            }
            if (lane == 0) {
              S.lpos[w][u][ln[u]] = cs;
              S.lmask[w][u][ln[u]] = make_uint2(am, 0u);
            }
            ln[u] += 1;
          }
        }
        if (open_any) {
          if (lane == 0) {
            snode[sp] = a_s;         // right
            snode[sp + 1] = nd + 1;  // left
#pragma unroll
            for (int u = 0; u < PPL; ++u) {
              S.smask[w][u][sp] = ms[u];
              S.smask[w][u][sp + 1] = ms[u];
            }
          }
          sp += 2;
        }
      }
    }
    __syncwarp();
  }
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
#pragma unroll
  for (int u = 0; u < PPL; ++u) {
    drain_list<EXACT, COUNTS, DW>(S.lpos[w][u], S.lmask[w][u], ln[u], lane, px[u], py[u], pz[u], ax[u], ay[u], az[u], cp[u]);
    if (slot[u] < slot_end) {
      if (!peer) {
        acc_t[3ull * slot[u] + 0] = ax[u];
        acc_t[3ull * slot[u] + 1] = ay[u];
        acc_t[3ull * slot[u] + 2] = az[u];
      } else {
        for (int r = 0; r < p2p.world; ++r) {
          double* dst = p2p.acc[r] + boff + 3ull * slot[u];
          dst[0] = ax[u];
          dst[1] = ay[u];
          dst[2] = az[u];
        }
      }
      if (COUNTS) {
        wcounts[4ull * slot[u] + 0] = cv[u];
        wcounts[4ull * slot[u] + 1] = ca[u];
        wcounts[4ull * slot[u] + 2] = cl[u];
        wcounts[4ull * slot[u] + 3] = cp[u];
      }
    }
  }
  if (PEER && peer) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t done = atomicAdd(p2p.cta_done, 1u);
      if (done == gridDim.x - 1) {
        *p2p.cta_done = 0;
        __threadfence_system();
        for (int r = 0; r < p2p.world; ++r) *reinterpret_cast<volatile uint32_t*>(p2p.flags[r] + p2p.rank) = epoch + 1u;
      }
    }
  }
}

}  // namespace kdnb
