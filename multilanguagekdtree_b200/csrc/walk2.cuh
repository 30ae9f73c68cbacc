// walk2.cuh — the walk kernel that runs by default (launched from walk.cu).
//
// theta-criterion force walk, calc_accel / accel_recur of the reference
// (Parallel/RustVersion/src/array_kd_tree.rs:585-621) with calc_pp_accel (array_particle.rs:67-76) in the leaves.
//
// One warp (= one CTA, 32 CTAs per SM) owns 32 consecutive TREE-ORDERED particles (a compact patch of ~4 leaves), one
// per lane, and keeps a shared-memory stack of frontier entries (node, lane mask).  Per round it pops up to 32
// entries and classifies them ONE NODE PER LANE against the bounding box of the 32 particles:
//   far   : size^2 <  theta^2 * dmin^2 * (1 - 1e-9)  -> every particle in the entry's mask accepts the node
//   near  : size^2 >= theta^2 * dmax^2 * (1 + 1e-9)  -> every particle opens it
//   mixed : otherwise -> the reference's test is evaluated per particle (below)
//   leaf  : its particles become pair interactions
// dmin/dmax are the distances from the node's centre of mass to the box; the 1e-9 margin dwarfs the <= 1e-15
// rounding of either side, so "far"/"near" provably agree with the reference's per-particle test
//       size*size < (THETA*THETA) * dist_sqr            (array_kd_tree.rs:606)
// Mixed nodes of a batch are parked in shared memory ({cm, size^2}, one slot per owning lane); the warp then runs over
// them with ONE PARTICLE PER LANE evaluating exactly that test with the reference's unfused operation order (:601-606),
// and a ballot hands the accept mask back to the owning lane.  After that the batch is finished one node per lane
// again: far + accepted monopoles are appended to the interaction list with their lane masks, near + still-open
// nodes push both children.  Every particle thus accepts / opens precisely the nodes the reference's recursion does
// (checked by KDNB_FLAG_WALK_COUNTS against the oracle: per-particle counts of tests, accepts, leaf visits and pair
// interactions are identical).  Leaves of a batch are queued and expanded 32/LP leaves at a time (LP lanes per leaf):
// lane -> (leaf, k) loads leaf particle k with one 32-byte access and appends it with the mask of the lanes that
// reached the leaf minus the owner (leaf_parts[i] != p, :590).
//
// Why this shape: on B200 an FP64 warp instruction holds its scheduler's issue port for two cycles and nothing
// else issues in its shadow (tools/issue_probe.cu: 16 DFMA = 35 cycles, every extra ALU instruction +1 cycle), so the
// kernel is bound by  2.19 * FP64 instructions + other instructions.  The first version of this kernel handled mixed
// nodes and leaves one at a time with shuffles (~90 issue cycles per mixed node, ~60 per leaf: 30 % of all
// instructions, profiles/README.md); here a mixed node costs 10 FP64 + ~11 other instructions and a leaf ~8.
//
// Forces are not evaluated during the traversal: the interaction list is drained by a branch-free, 4-way unrolled
// loop, 16 FP64 + 7 other instructions per interaction (broadcast shared-memory loads, no global loads).  When every
// z coordinate is +-0 and every mass is > 0 (flat[3], set by flat_detect in sort.cu — the reference's own initial
// conditions, circular_orbits, array_particle.rs:19-44, are planar for ever) all centre-of-mass z are +-0 too, dz is
// exactly 0 in every interaction and every test, and the z terms are skipped (13 FP64 per interaction): the
// results are bit-identical to the general path, az stays +0.
//
// Accumulation is a running f64 sum per particle (the reference combines pairwise along the recursion, :611-613);
// the difference is summation order only and is covered by the stated 1e-12 tolerance.
#pragma once
#include "walk_legacy.cuh"

namespace kdnb {

constexpr int W2_STACK = 320;  // soft capacity: batches shrink as the stack fills
constexpr int W2_SLACK = 32;   // depth-first tail when the stack is at capacity (tree depth <= 27 at 1e8 particles)
constexpr int W2_LIST = 96;    // interaction-list capacity (appends come in groups of <= 32)

// 1.875 = the e^2 coefficient of (1 - e)^(-3/2); read from the constant bank as an instruction operand (as a literal
// it costs two register moves per loop iteration at the 64-register budget)
__constant__ double W2_C2 = 1.875;

template <bool EXACT>
struct W2Smem {
  uint32_t snode[W2_STACK + W2_SLACK];
  uint32_t smask[W2_STACK + W2_SLACK];
  double lx[W2_LIST], ly[W2_LIST], lz[W2_LIST], lm[W2_LIST];  // monopoles / leaf particles, SoA: the drain loop reads
  uint32_t lmask[W2_LIST];                                     // four consecutive interactions with 16-byte loads
  uint32_t lflag[EXACT ? W2_LIST : 4];  // 1 = leaf particle (the exact-math formulas differ, see interact<>)
  union {
    Rec32 mix[32];  // {cx, cy, cz, size^2} of the batch's mixed nodes, slot = owning lane
    uint4 lq[32];   // {first slot, num_parts, lane mask, -} of the batch's leaves, compacted
  };
};

// all lanes stream over the list; lane = particle
template <bool EXACT, bool FLATZ>
__device__ __forceinline__ void drain2(W2Smem<EXACT>& S, int cnt, int lane, double px, double py, double pz, double& ax,
                                       double& ay, double& az) {
  __syncwarp();
  if (EXACT) {
    for (int i = 0; i < cnt; ++i) {
      Rec32 e;
      e.a = S.lx[i], e.b = S.ly[i], e.c = S.lz[i], e.d = S.lm[i];
      const bool use = (S.lmask[i] >> lane) & 1u;
      interact<true>(e, use, S.lflag[i] != 0, px, py, pz, ax, ay, az);
    }
  } else {
    constexpr int DW = 4;
    // DW interactions in lock-step, interleaved stage by stage; the list is padded to a multiple of DW with
    // masked-out entries.  A masked-out lane zeroes the rsqrt estimate (one 32-bit select: MUFU.RSQ64H leaves the low
    // word 0), which makes its contribution exactly -0 * d = no-op and also absorbs d2 == 0 (inf estimate).
    const uint32_t lanebit = 1u << lane;
    const int padded = (cnt + DW - 1) / DW * DW;
    if (lane < padded - cnt) {
      S.lx[cnt + lane] = S.ly[cnt + lane] = S.lz[cnt + lane] = S.lm[cnt + lane] = 0.0;
      S.lmask[cnt + lane] = 0u;
    }
    __syncwarp();
    for (int i = 0; i < padded; i += DW) {
      double dx[DW], dy[DW], dz[DW], d2[DW], y[DW], y2[DW], ee[DW], mq[DW], q[DW];
      const uint4 m4 = *reinterpret_cast<const uint4*>(&S.lmask[i]);
      const uint32_t use[4] = {m4.x & lanebit, m4.y & lanebit, m4.z & lanebit, m4.w & lanebit};
#pragma unroll
      for (int j = 0; j < DW; j += 2) {
        const double2 ex = *reinterpret_cast<const double2*>(&S.lx[i + j]);
        const double2 ey = *reinterpret_cast<const double2*>(&S.ly[i + j]);
        const double2 em = *reinterpret_cast<const double2*>(&S.lm[i + j]);
        dx[j] = __dsub_rn(px, ex.x), dx[j + 1] = __dsub_rn(px, ex.y);
        dy[j] = __dsub_rn(py, ey.x), dy[j + 1] = __dsub_rn(py, ey.y);
        if (!FLATZ) {
          const double2 ez = *reinterpret_cast<const double2*>(&S.lz[i + j]);
          dz[j] = __dsub_rn(pz, ez.x), dz[j + 1] = __dsub_rn(pz, ez.y);
        }
        mq[j] = -em.x, mq[j + 1] = -em.y;
      }
#pragma unroll
      for (int j = 0; j < DW; ++j) d2[j] = __dmul_rn(dx[j], dx[j]);
#pragma unroll
      for (int j = 0; j < DW; ++j) d2[j] = fma(dy[j], dy[j], d2[j]);
      if (!FLATZ) {
#pragma unroll
        for (int j = 0; j < DW; ++j) d2[j] = fma(dz[j], dz[j], d2[j]);
      }
#pragma unroll
      for (int j = 0; j < DW; ++j) {
        const double r = rsqrt_estimate(d2[j]);
        y[j] = __hiloint2double(use[j] ? __double2hiint(r) : 0, 0);
      }
#pragma unroll
      for (int j = 0; j < DW; ++j) y2[j] = __dmul_rn(y[j], y[j]);
#pragma unroll
      for (int j = 0; j < DW; ++j) {
        ee[j] = fma(-d2[j], y2[j], 1.0);
        y[j] = __dmul_rn(y[j], y2[j]);  // y0^3
      }
#pragma unroll
      for (int j = 0; j < DW; ++j) {
        q[j] = fma(W2_C2, ee[j], 1.5);
        mq[j] = __dmul_rn(mq[j], y[j]);  // -m * y0^3
      }
#pragma unroll
      for (int j = 0; j < DW; ++j) ee[j] = __dmul_rn(mq[j], ee[j]);
#pragma unroll
      for (int j = 0; j < DW; ++j) mq[j] = fma(ee[j], q[j], mq[j]);
#pragma unroll
      for (int j = 0; j < DW; ++j) {
        ax = fma(mq[j], dx[j], ax);
        ay = fma(mq[j], dy[j], ay);
        if (!FLATZ) az = fma(mq[j], dz[j], az);
      }
    }
  }
  __syncwarp();
}

enum : int { W2_NONE = 0, W2_FAR = 1, W2_NEAR = 2, W2_MIXED = 3, W2_LEAF = 4 };

template <bool EXACT, bool COUNTS, bool PEER, bool FLATZ>
__device__ __forceinline__ void walk2_body(W2Smem<EXACT>& S, const WNode* __restrict__ nodes,
                                           const PosM* __restrict__ posm, double* __restrict__ acc_t,
                                           uint32_t slot_begin, uint32_t slot_end, double theta2,
                                           unsigned long long* __restrict__ wcounts, const P2P& p2p, int lshift,
                                           const uint32_t* __restrict__ gorder, uint32_t* __restrict__ gcost) {
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
#pragma unroll
  for (int k = 0; k < ND; ++k) {
    lo[k] = warp_min(lo[k]);
    hi[k] = warp_max(hi[k]);
  }
  {
    const uint32_t m0 = __ballot_sync(0xffffffffu, valid_p);
    if (lane == 0) {
      S.snode[0] = 0;
      S.smask[0] = m0;
    }
  }
  __syncwarp();
  const double far_margin = 1.0 - 1e-9, near_margin = 1.0 + 1e-9;
  const int kmask = (1 << lshift) - 1, lpr = 32 >> lshift;

  int ln = 0;
  int sp = warp_has_work ? 1 : 0;
  while (sp > 0) {
    // ---- pop a batch: lane l takes entry sp+l after the pop (order inside a batch is irrelevant)
    const int room = W2_STACK - sp;
    const int nb = min(min(sp, 32), max(1, room));
    sp -= nb;
    work += 4u;
    const bool has = lane < nb;
    uint32_t node = 0, na = 0, nbits = 0, mk = 0;
    int kind = W2_NONE;
    Rec32 c;
    c.a = c.b = c.c = c.d = 0.0;
    double size2 = 0.0;
    if (has) {
      node = S.snode[sp + lane];
      mk = S.smask[sp + lane];
      const Rec32* rec = reinterpret_cast<const Rec32*>(nodes + node);
      const int4 info = __ldg(reinterpret_cast<const int4*>(rec + 1));  // size2, (a, b)
      c = rec[0];  // cx, cy, cz, m (unused for a leaf; loaded alongside so the two sectors travel together)
      na = (uint32_t)info.z;
      nbits = (uint32_t)info.w;
      kind = W2_LEAF;
      if (nbits & WN_INTERNAL) {
        size2 = __hiloint2double(info.y, info.x);
        double dmin2 = 0.0, dmax2 = 0.0;
        const double cc[3] = {c.a, c.b, c.c};
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          const double below = lo[k] - cc[k], above = cc[k] - hi[k];  // > 0 when the centre is outside the box
          const double dn = fmax(0.0, fmax(below, above));
          const double df = fmax(fabs(below), fabs(above));  // = max(|c-lo|, |c-hi|)
          dmin2 = fma(dn, dn, dmin2);
          dmax2 = fma(df, df, dmax2);
        }
        kind = W2_MIXED;
        if (size2 < theta2 * dmin2 * far_margin) kind = W2_FAR;
        else if (size2 >= theta2 * dmax2 * near_margin) kind = W2_NEAR;
      }
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
      if (kind == W2_MIXED) {
        Rec32 t;
        t.a = c.a, t.b = c.b, t.c = c.c, t.d = size2;
        S.mix[lane] = t;
      }
      __syncwarp();
      uint32_t am = 0;
      for (uint32_t rem = bal_mixed; rem; rem &= rem - 1) {
        const int k = __ffs(rem) - 1;
        const Rec32 q = S.mix[k];
        const double dx = __dsub_rn(px, q.a), dy = __dsub_rn(py, q.b);
        double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));  // :604, left to right
        if (!FLATZ) {
          const double dz = __dsub_rn(pz, q.c);
          d2 = __dadd_rn(d2, __dmul_rn(dz, dz));
        }
        const bool accept = q.d < __dmul_rn(theta2, d2);  // :606
        const uint32_t b = __ballot_sync(0xffffffffu, accept);
        if (lane == k) am = b;
        if (COUNTS) {
          const uint32_t mk_k = __shfl_sync(0xffffffffu, mk, k);
          const bool in = (mk_k >> lane) & 1u;
          cv += in;
          ca += in && accept;
        }
      }
      if (kind == W2_MIXED) {
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
        if (ln + add > W2_LIST) {
          drain2<EXACT, FLATZ>(S, ln, lane, px, py, pz, ax, ay, az);
          work += (uint32_t)ln;
          ln = 0;
        }
        if (mine) {
          const int i = ln + __popc(bal & lt);
          S.lx[i] = c.a, S.ly[i] = c.b, S.lz[i] = c.c, S.lm[i] = c.d;
          S.lmask[i] = amask;
          if (EXACT) S.lflag[i] = 0u;
        }
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
        PosM qv;
        if (valid) qv = posm[j];
        if (ln + add > W2_LIST) {
          drain2<EXACT, FLATZ>(S, ln, lane, px, py, pz, ax, ay, az);
          work += (uint32_t)ln;
          ln = 0;
        }
        if (valid) {
          const int i = ln + __popc(bal & lt);
          S.lx[i] = qv.x, S.ly[i] = qv.y, S.lz[i] = qv.z, S.lm[i] = qv.m;
          S.lmask[i] = m;
          if (EXACT) S.lflag[i] = 1u;
        }
        ln += add;
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
  drain2<EXACT, FLATZ>(S, ln, lane, px, py, pz, ax, ay, az);
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
             uint32_t* __restrict__ gcost) {
  pdl_sync();
  __shared__ W2Smem<EXACT> S;
  if (!EXACT && !COUNTS && flat[3])
    walk2_body<EXACT, COUNTS, PEER, true>(S, nodes, posm, acc_t, slot_begin, slot_end, theta2, wcounts, p2p, lshift, gorder, gcost);
  else
    walk2_body<EXACT, COUNTS, PEER, false>(S, nodes, posm, acc_t, slot_begin, slot_end, theta2, wcounts, p2p, lshift, gorder, gcost);
}

// Heaviest-first launch order for the next walk (one CTA): gorder = the groups sorted by descending gcost, by a
// counting sort on 1024 cost classes.  Every walk launch takes ~0.2 ms beyond its issue-slot work whatever the grid
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
