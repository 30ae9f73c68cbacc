// ctx.cuh — the context behind the opaque kdnb_ctx handle: device buffers, level plan, stream, counters.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"

namespace kdnb {

constexpr int SORT_THREADS = 256;
constexpr int SORT_IPT = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;  // keys per CTA
constexpr int LVL_THREADS = 256;
constexpr int LVL_CHUNK = 2048;  // list entries per CTA in the global-level partition kernels

// Peer-memory exchange (multi-GPU): every rank's walk kernel stores its accelerations straight into all peers'
// acc_t buffers over NVLink (cudaIpc-mapped), then raises a per-rank flag on every peer; a one-CTA wait kernel on each
// GPU spins on its local flags.  Two acc buffers alternate by step parity so a fast rank never overwrites a buffer a
// slow rank is still consuming.
constexpr int P2P_MAX = 16;
struct P2P {
  double* acc[P2P_MAX];      // peers' acc_t bases, own included
  uint32_t* flags[P2P_MAX];  // peers' flag arrays [P2P_MAX]
  const uint32_t* epoch;     // local step counter (device)
  uint32_t* cta_done;        // local "CTAs finished" counter (device)
  uint64_t stride;           // doubles between the two acc buffers
  int world, rank;
};

// Sharded tree build (multi-GPU, peer mode, world = 2^k): the top k levels are built by every rank, below them rank r
// builds only the subtree of level-k segment r and stores its node records, its slice of the tree order (perm) and its
// root's mass sums straight into every peer's buffers; a flag per rank publishes them (build.cu).
// Split sort (multi-GPU, peer mode): a rank sorts only every m-th non-flat dimension (m = min(world, dimensions)), leaves
// its sorted lists in an export buffer and fetches the others from a peer that sorted them (sort.cu).
// p2p_state words: [0] acc epoch, [1] acc cta_done, [2] error, [4..20) acc flags, [20..36) build flags, [36] build epoch,
// [37] build cta_done, [38..54) sort flags, [54] sort epoch, [55] export cta_done, [56] fetch cta_done
constexpr int P2P_BFLAGS = 4 + P2P_MAX, P2P_BEPOCH = 4 + 2 * P2P_MAX, P2P_BDONE = 5 + 2 * P2P_MAX;
constexpr int P2P_SFLAGS = 6 + 2 * P2P_MAX, P2P_SEPOCH = 6 + 3 * P2P_MAX, P2P_SDONE = 7 + 3 * P2P_MAX, P2P_FDONE = 8 + 3 * P2P_MAX,
              P2P_WORDS = 12 + 3 * P2P_MAX;
struct P2PBuild {
  WNode* nodes[P2P_MAX];     // peers' node arrays, own included
  uint32_t* perm[P2P_MAX];   // peers' tree order
  double4* ms[P2P_MAX];      // peers' per-node mass sums
  uint32_t* state[P2P_MAX];  // peers' p2p_state
  uint32_t* sexp[P2P_MAX];   // peers' sorted-list export buffers ([3][n] ids; their keys[1], dead between two sorts)
  int world, rank;
};

struct Ctx {
  // configuration
  int device = 0;
  int num_sms = 148;
  uint32_t mp = 8;
  int layout = KDNB_LAYOUT_PADDED;
  double theta = 0.3, theta2 = 0.09;
  uint32_t flags = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side_stream = nullptr;  // captures conditional-node bodies (sort.cu)
  cudaStream_t order_stream = nullptr; // the walk's launch-order kernel, next to the build (walk.cu)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool order_pending = false;

  // sizes
  uint64_t n = 0;        // particles (N+1 of the reference)
  bool empty = false;    // an empty particle set was uploaded: every stage is a no-op, as in the reference
  uint64_t cap = 0;      // allocated particle capacity
  uint64_t n_nodes = 0;  // allocate_node_vec length for (n, layout)
  uint64_t node_cap = 0;
  int l0 = 0;            // first level whose segments fit the bottom kernel (<= BOT_CAP)
  uint32_t ntiles = 0;   // sort tiles

  // particle state, SoA, original order (Particle, array_particle.rs:3-8)
  double* pos[3] = {nullptr, nullptr, nullptr};
  double* vel[3] = {nullptr, nullptr, nullptr};
  double* mass = nullptr;
  double* radius = nullptr;
  kdnb_particle* aos = nullptr;  // staging for AoS <-> SoA conversion
  kdnb_particle_simd* aos_simd = nullptr;  // staging of the 96-byte SIMD records (allocated by the first kdnb_*_simd call)
  uint64_t aos_simd_cap = 0;
  PosM* pm = nullptr;            // {x, y, z, m} in original order: one-sector gathers for the bottom build kernel

  // build scratch
  uint64_t* keys[2] = {nullptr, nullptr};  // [3][n] radix keys, ping-pong
  uint32_t* list[2] = {nullptr, nullptr};  // [3][n] per-dimension sorted id lists, ping-pong
  uint32_t* hist = nullptr;                // [3][256][ntiles]
  uint32_t* digit_tot = nullptr;           // [3][256]
  uint32_t* flat = nullptr;                // [3] 1 = all coordinates of that dimension equal (lives behind digit_tot); [3] planar walk
  uint32_t* dmask = nullptr;               // [0..2] 1 = this rank does not sort dimension d (flat, or a peer's share of the split sort); [4..6] the peer to fetch it from (behind flat)
  uint64_t* sort_state = nullptr;          // [SS_WORDS] 32-bit key scaling, need64 flag, running position extents (common.cuh)
  bool extent_fresh = false;               // the extent records describe the current positions and were not consumed yet
  uint32_t* rk = nullptr;                  // [3][n] rank of every particle in the initial sorted list of each dimension
  uint4* tseg = nullptr;                   // level table, level l at offset 2^l - 1: {first slot, length, node, buffer bits}
  uint32_t* inv = nullptr;                 // [n] id -> local slot inside a bottom segment
  uint64_t* lvl_status = nullptr;          // [3][nseg][chunks] look-back status words of level_partition (build.cu)
  uint32_t* lvl_ctl = nullptr;             // [72] per-level tickets, build epoch, timeout flag
  uint64_t table_cap = 0, chunk_cap = 0;

  // tree + tree-ordered views
  WNode* nodes = nullptr;    // reference node layout (padded or dense indices)
  double4* ms = nullptr;     // per node {M, sum m*x, sum m*y, sum m*z}
  uint32_t* perm = nullptr;  // tree slot -> particle id (the reference's `indices` after the build)
  uint32_t* rank = nullptr;  // particle id -> tree slot
  PosM* posm = nullptr;      // tree-ordered {x,y,z,m}
  double* acc_t = nullptr;   // [slots_pad][3] tree-ordered accelerations (the exchanged array)
  unsigned long long* wcounts = nullptr;  // [n][4] optional walk counters, tree order
  double* tmp3 = nullptr;    // [n][3] staging for accel up/download
  uint32_t* gcost = nullptr;   // [groups] work of every 32-slot group in the last production walk (walk.cu)
  uint32_t* gorder = nullptr;  // [groups] launch order of the next walk: heaviest group first
  uint64_t gcost_n = 0;        // particle count / shard the recorded work belongs to (0: none)
  uint32_t* wseed = nullptr;   // [64] node indices of tree depths 0-5 in heap order (walk2.cuh: the groups' common descent), plan()
  bool wseed_ok = false;       // every node of depths 0-4 is internal for this (n, max_parts)
  uint64_t wseed_n = ~0ull;    // particle count the table was computed for
  uint32_t gcost_begin = 0, gcost_end = 0;
  bool tree_valid = false, acc_valid = false, map_valid = false;
  bool bottom_attr_set = false;  // build_bottom's dynamic shared memory limit raised on this context's device
  uint64_t planned_n = 0;

  // multi-GPU
  int rank_id = 0, world = 1;
  void* nccl_comm = nullptr;
  uint64_t shard_slots = 0;  // tree slots per rank (multiple of 32)
  bool use_pdl = false;      // programmatic dependent launch between the kernels of a step (opt-in: KDNB_PDL=1)
  bool p2p_ready = false;    // peer set-up attempted for the current allocation
  bool p2p_on = false;       // accelerations exchanged by peer stores inside the walk kernel (else ncclAllGather)
  P2P p2p = {};
  uint32_t* p2p_state = nullptr;  // device: P2P_WORDS words, layout above
  void* p2p_mapped[6 * P2P_MAX] = {};
  P2PBuild p2pb = {};
  bool split_sort = false;   // the last sort_prep assigned the dimensions for a split sort
  int shard_k = 0;           // > 0: the build below level shard_k is sharded over the 2^shard_k ranks (this step)
  uint64_t acc_stride = 0;   // doubles per acc buffer

  // one step captured as a CUDA graph (replayed by kdnb_simple_sim when not profiling)
  cudaGraphExec_t step_graph = nullptr;
  uint64_t graph_n = 0, graph_launches = 0;
  double graph_dt = 0.0;
  int graph_world = 0;
  uint64_t seen_n = 0, seen_calls = 0;  // consecutive kdnb_simple_sim calls with the same (n, dt)
  double seen_dt = 0.0;

  // measurement
  uint64_t launches = 0;
  cudaEvent_t pev[5] = {};      // KDNB_FLAG_PROFILE: stage boundaries of the current step (event-record nodes in the step graph)
  double stage_acc[4] = {0, 0, 0, 0};  // build / walk / exchange / kick ms summed since the last reset
  uint64_t ev_steps = 0;
  cudaEvent_t sw_begin = nullptr, sw_end = nullptr;
  void* l2_scratch = nullptr;

  mutable std::string err;
  int fail(int code, const std::string& msg) const {
    err = msg;
    return code;
  }
};

// sort.cu
int sort_lists(Ctx* c);
// build.cu
int build_tree(Ctx* c);
int export_tree(Ctx* c, kdnb_node* dev_out);
void init_unused_nodes(Ctx* c);
void subtree_of(uint64_t n, uint32_t mp, int layout, int k, uint32_t seg, uint64_t* a, uint64_t* len, uint64_t* node);
// walk.cu
int walk(Ctx* c);
void walk_order_fork(Ctx* c);
// kick.cu
int kick_drift(Ctx* c, double dt);
int p2p_wait_step(Ctx* c);
int aos_to_soa(Ctx* c);
int simd_to_soa(Ctx* c, const kdnb_particle_simd* dev_aos);
int soa_to_simd(Ctx* c, kdnb_particle_simd* dev_aos);
int soa_to_aos(Ctx* c);
int gather_acc(Ctx* c, double* dst_orig_order);
int scatter_acc(Ctx* c, const double* src_orig_order);
int gather_counts(Ctx* c, unsigned long long* dst_orig_order);
// select.cu
int select_run(Ctx* c, const double* d_vals, const uint64_t* d_idx64, uint64_t count, uint64_t goal, uint64_t* d_keys,
               uint32_t* d_idx32, uint2* d_tiles, uint64_t* d_state, uint64_t* d_out64);
// peak.cu
int measure_fp64_peak(Ctx* c, double* tflops);

// Programmatic dependent launch (opt-in, KDNB_PDL=1): every kernel of the step can be launched with the
// programmatic-stream-serialization attribute and starts with pdl_sync() — it signals that ITS successor may be scheduled (the successor's launch and
// prologue then overlap this kernel's tail) and waits until its PREDECESSOR has completed and flushed its writes.
// The step is a chain of ~50 kernels of 5-35 us at N=1M, so the launch gaps are a visible share of the build.
__device__ __forceinline__ void pdl_sync() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

template <typename... KArgs, typename... Args>
inline void kdnb_launch(Ctx* c, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = c->use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
  c->launches++;
}

#define KDNB_LAUNCH(ctx, kernel, grid, block, smem, ...) kdnb_launch((ctx), kernel, dim3(grid), dim3(block), (smem), __VA_ARGS__)

#define KDNB_CHECK_LAUNCH(ctx)                                                   \
  do {                                                                           \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) return (ctx)->fail(KDNB_E_CUDA, std::string(__FILE__) + ": " + cudaGetErrorString(_e)); \
  } while (0)

}  // namespace kdnb
