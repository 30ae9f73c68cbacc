// walk.cu — theta-criterion force walk, calc_accel / accel_recur of the reference
// (Parallel/RustVersion/src/array_kd_tree.rs:585-621) with calc_pp_accel (array_particle.rs:67-76) in the leaves:
// host-side launch logic, the peer-exchange wait kernel, and the kernels themselves (walk2.cuh).
#include <cstdlib>

#include "ctx.cuh"
#include "walk2.cuh"

namespace kdnb {

// one CTA per GPU: wait until every rank (this one included) has published the accelerations of the current step
__global__ void p2p_wait_kernel(uint32_t* state, int world) {
  pdl_sync();
  const uint32_t target = state[0] + 1u;
  volatile uint32_t* flags = state + 4;
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
  if (threadIdx.x == 0) state[0] = target;
}

int p2p_wait_step(Ctx* c) {
  KDNB_LAUNCH(c, p2p_wait_kernel, 1, 32, 0, c->p2p_state, c->world);
  KDNB_CHECK_LAUNCH(c);
  return 0;
}

// this rank's share of the tree slots
static void walk_range(const Ctx* c, uint32_t* begin, uint32_t* end) {
  *begin = 0, *end = (uint32_t)c->n;
  if (c->world > 1) {
    *begin = (uint32_t)std::min<uint64_t>(c->n, (uint64_t)c->rank_id * c->shard_slots);
    *end = (uint32_t)std::min<uint64_t>(c->n, (uint64_t)(c->rank_id + 1) * c->shard_slots);
  }
}

// Heaviest-first launch order from the work the groups reported in the previous walk of the same particle set and
// shard (the tree order moves little from step to step).  Production kernel only, and only while the node records
// and tree-ordered particles (~66 B per particle) fit the 126 MB L2: the index order is also the spatial order, so
// the ~3500 groups in flight share their near field; with the cost order they are scattered over the whole domain,
// which costs nothing when the tree is L2-resident (measured at N = 125k and 1M).  Where it is not, what counts is the
// length of the launch: one GPU walking all of N = 10M loses (20.69 ms in index order, 20.99 ms heaviest first,
// profiles/r02_ab_walk_lpt_10M.txt), a 1/8 shard of the same tree wins (2.69 -> 2.56 ms on 8 GPUs,
// profiles/r02_scale_8gpu.txt) because the end of the launch it shortens is a larger part of it.  So: heaviest first
// while this rank walks at most 2M slots.  KDNB_WALK_LPT=0 disables, =1 forces it at every size.
static bool walk_uses_order(const Ctx* c, uint32_t begin, uint32_t end) {
  static const int lpt_mode = [] {
    const char* s = getenv("KDNB_WALK_LPT");
    return s ? (atoi(s) != 0 ? 1 : 0) : -1;
  }();
  constexpr uint64_t LPT_MAX_N = 1ull << 21;
  const bool lpt = lpt_mode == 1 || (lpt_mode < 0 && (uint64_t)(end - begin) <= LPT_MAX_N);
  const bool production = !(c->flags & (KDNB_FLAG_WALK_COUNTS | KDNB_FLAG_EXACT_MATH));
  return production && lpt && end > begin;
}
static bool walk_has_costs(const Ctx* c, uint32_t begin, uint32_t end) {
  return c->gcost_n == c->n && c->gcost_begin == begin && c->gcost_end == end && (end - begin + 31) / 32 > 1;
}

// The order kernel (one CTA, ~27 us at N = 1M) depends only on the previous walk: a step starts it on a side stream
// next to the tree build and the walk waits for it (a fork / join inside the step graph).
void walk_order_fork(Ctx* c) {
  uint32_t begin, end;
  walk_range(c, &begin, &end);
  c->order_pending = false;
  if (!c->order_stream || !walk_uses_order(c, begin, end) || !walk_has_costs(c, begin, end)) return;
  if (cudaEventRecord(c->ev_fork, c->stream) != cudaSuccess || cudaStreamWaitEvent(c->order_stream, c->ev_fork, 0) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  cudaStream_t keep = c->stream;
  c->stream = c->order_stream;
  KDNB_LAUNCH(c, walk_order_kernel, 1, 1024, 0, c->gcost, c->gorder, (end - begin + 31) / 32);
  c->stream = keep;
  cudaEventRecord(c->ev_join, c->order_stream);
  c->order_pending = true;
}

// grids of at least this many waves of 28 CTAs per SM use the 72-register build of the production kernel
#ifndef KDNB_W2_MINB28_WAVES
#define KDNB_W2_MINB28_WAVES 2
#endif
constexpr uint32_t W2_MINB28_WAVES = KDNB_W2_MINB28_WAVES;

static void launch_walk2(Ctx* c, uint32_t begin, uint32_t end) {
  const uint32_t grid = std::max<uint32_t>((end - begin + 31) / 32, 1u);
  const bool counts = (c->flags & KDNB_FLAG_WALK_COUNTS) != 0;
  const bool exact = (c->flags & KDNB_FLAG_EXACT_MATH) != 0;
  P2P pp = c->p2p;
  if (!c->p2p_on) pp.world = 0;
  int lshift = 0;  // lanes per leaf in the leaf rounds: the smallest power of two >= MAX_PARTS
  while ((1u << lshift) < c->mp) ++lshift;
  const uint32_t* gorder = nullptr;
  uint32_t* gcost = nullptr;
  if (walk_uses_order(c, begin, end)) {
    if (c->order_pending) {  // started by walk_order_fork at the beginning of this step
      cudaStreamWaitEvent(c->stream, c->ev_join, 0);
      gorder = c->gorder;
    } else if (walk_has_costs(c, begin, end)) {
      KDNB_LAUNCH(c, walk_order_kernel, 1, 1024, 0, c->gcost, c->gorder, grid);
      gorder = c->gorder;
    }
    gcost = c->gcost;
    c->gcost_n = c->n;
    c->gcost_begin = begin;
    c->gcost_end = end;
  }
  c->order_pending = false;
#define KDNB_WALK_ARGS c->nodes, c->posm, c->acc_t, begin, end, c->theta2, c->wcounts, pp, c->flat, lshift, gorder, gcost, (c->wseed_ok ? c->wseed : nullptr)
  const bool peer = pp.world > 1;
  if (exact && counts)
    KDNB_LAUNCH(c, (walk2_kernel<true, true, true, 1>), grid, 32, 0, KDNB_WALK_ARGS);
  else if (exact)
    KDNB_LAUNCH(c, (walk2_kernel<true, false, true, 1>), grid, 32, 0, KDNB_WALK_ARGS);
  else if (counts)
    KDNB_LAUNCH(c, (walk2_kernel<false, true, true, 1>), grid, 32, 0, KDNB_WALK_ARGS);
  else {
    // Production kernel.  Register budget by grid size: 24 one-warp CTAs per SM = 80 registers (25 resident), or 28 per
    // SM = 72 registers (40 bytes spilled; the 7296 bytes of shared memory are sized so that 28 fit).  More resident
    // warps win once the grid is several waves deep — 2.046 -> 2.025 ms at N=1M (31251 CTAs), 20.15 -> 19.69 ms at
    // N=10M — and lose on a grid of about one wave, where every CTA is resident either way and only the spills remain
    // (0.289 -> 0.303 ms at 3907 CTAs, the 1/8 shard of N=1M); profiles/r02_ab_walk_unroll.txt, r02_ab_walk_minb_grid.txt.
    // 32 per SM at 64 registers is slower at every size (profiles/r01_ab_walk_minb.txt); nor is the time a launch takes
    // beyond its issue-slot work a matter of node / leaf load latency: prefetching the right child at push time and a
    // leaf's particles at classification made every size ~1 % slower (profiles/r01_ab_walk_prefetch.txt).
    static const int minb_env = [] {
      const char* s = getenv("KDNB_WALK_MINB");  // profiling knob: 24 / 28 / 32 at every grid size
      return s ? atoi(s) : 0;
    }();
    const int minb = minb_env ? minb_env : (grid >= W2_MINB28_WAVES * 28u * (uint32_t)c->num_sms ? 28 : 24);
#ifdef KDNB_WALK_AB  // development experiment: KDNB_WALK_PAD=<bytes> of unused dynamic shared memory limits the CTAs per SM
    static const int pad = [] { const char* s = getenv("KDNB_WALK_PAD"); return s ? atoi(s) : 0; }();
#else
    constexpr int pad = 0;
#endif
    if (peer) {
      if (minb == 28) KDNB_LAUNCH(c, (walk2_kernel<false, false, true, 28>), grid, 32, 0, KDNB_WALK_ARGS);
      else KDNB_LAUNCH(c, (walk2_kernel<false, false, true, 24>), grid, 32, 0, KDNB_WALK_ARGS);
    } else if (minb == 32) KDNB_LAUNCH(c, (walk2_kernel<false, false, false, 32>), grid, 32, 0, KDNB_WALK_ARGS);
    else if (minb == 28) KDNB_LAUNCH(c, (walk2_kernel<false, false, false, 28>), grid, 32, 0, KDNB_WALK_ARGS);
    else KDNB_LAUNCH(c, (walk2_kernel<false, false, false, 24>), grid, 32, pad, KDNB_WALK_ARGS);
  }
#undef KDNB_WALK_ARGS
}

int walk(Ctx* c) {
  uint32_t begin, end;
  walk_range(c, &begin, &end);
  // (peer mode: a rank whose shard is empty — fewer particles than 64 x (world - 1) — still launches one CTA, which has
  // nothing to walk but raises this rank's flag on every peer; without it all ranks would wait for the flag until the
  // wait kernel's timeout, every step)
  if (end > begin || (c->world > 1 && c->p2p_on)) {
#ifdef KDNB_WALK_AB  // development experiments only (-DKDNB_WALK_AB): KDNB_WALK_DBG=1 skips the drains, =2 doubles them
    {
      static const int dbg = [] { const char* s = getenv("KDNB_WALK_DBG"); return s ? atoi(s) : 0; }();
      static bool set = false;
      if (!set) cudaMemcpyToSymbol(w2_dbg, &dbg, sizeof(int)), set = true;
    }
#endif
    launch_walk2(c, begin, end);
    KDNB_CHECK_LAUNCH(c);
  }
  c->acc_valid = true;
  return 0;
}

}  // namespace kdnb
