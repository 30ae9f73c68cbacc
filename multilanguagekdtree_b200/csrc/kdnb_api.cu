// kdnb_api.cu — the C ABI of libkdnb.so (include/kdnb.h): context, transfers, the step driver
// (simple_sim, Parallel/RustVersion/src/array_kd_tree.rs:623-664), NCCL exchange and measurement hooks.
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "ctx.cuh"

using namespace kdnb;

struct kdnb_ctx {
  Ctx c;
};

static thread_local std::string g_create_error;

// ------------------------------------------------------------------------------------------------ NCCL (dlopen)
namespace {
struct NcclId {
  char internal[128];
};
typedef int (*fn_get_unique_id)(NcclId*);
typedef int (*fn_comm_init_rank)(void**, int, NcclId, int);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_comm_destroy)(void*);
typedef const char* (*fn_get_error_string)(int);
struct NcclApi {
  void* handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_get_error_string get_error_string = nullptr;
  std::once_flag once;
} g_nccl;

bool load_nccl(std::string* why) {
  std::call_once(g_nccl.once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.handle) break;
    }
    if (g_nccl.handle) {
      g_nccl.get_unique_id = (fn_get_unique_id)dlsym(g_nccl.handle, "ncclGetUniqueId");
      g_nccl.comm_init_rank = (fn_comm_init_rank)dlsym(g_nccl.handle, "ncclCommInitRank");
      g_nccl.all_gather = (fn_all_gather)dlsym(g_nccl.handle, "ncclAllGather");
      g_nccl.comm_destroy = (fn_comm_destroy)dlsym(g_nccl.handle, "ncclCommDestroy");
      g_nccl.get_error_string = (fn_get_error_string)dlsym(g_nccl.handle, "ncclGetErrorString");
      if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.all_gather || !g_nccl.comm_destroy) {
        dlclose(g_nccl.handle);
        g_nccl.handle = nullptr;
      }
    }
  });
  if (!g_nccl.handle && why) *why = "libnccl.so.2 could not be loaded";
  return g_nccl.handle != nullptr;
}
constexpr int KDNB_MAX_WORLD = 64;  // ranks of one communicator (kdnb_comm_init)
constexpr int NCCL_FLOAT64 = 8;  // ncclDouble
constexpr int NCCL_UINT8 = 1;
}  // namespace

// ------------------------------------------------------------------------------------------------ helpers

template <typename T>
static int dev_alloc(Ctx* c, T** p, uint64_t count) {
  if (*p) {
    cudaFree(*p);
    *p = nullptr;
  }
  if (count == 0) count = 1;
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e != cudaSuccess) {
    *p = nullptr;
    return c->fail(KDNB_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  return 0;
}

static void close_peers(Ctx* c);
static void peer_barrier(Ctx* c);

static void free_all(Ctx* c) {
  auto fr = [](auto*& p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  // acc_t, p2p_state, nodes, perm, ms and keys[1] were exported with cudaIpcGetMemHandle: every rank closes its mappings of the
  // peers' buffers and all ranks meet before anyone frees (freeing exported memory an importer still maps is undefined
  // behaviour).  free_all runs from a growing plan() and from kdnb_destroy, both collective calls when world > 1.
  const bool exported = c->p2p_on;
  close_peers(c);
  if (exported) peer_barrier(c);
  for (int d = 0; d < 3; ++d) {
    fr(c->pos[d]);
    fr(c->vel[d]);
  }
  fr(c->mass);
  fr(c->radius);
  fr(c->aos);
  fr(c->aos_simd);
  c->aos_simd_cap = 0;
  fr(c->keys[0]);
  fr(c->keys[1]);
  fr(c->list[0]);
  fr(c->list[1]);
  fr(c->hist);
  fr(c->digit_tot);
  fr(c->sort_state);
  fr(c->rk);
  fr(c->pm);
  fr(c->inv);
  fr(c->tseg);
  fr(c->lvl_status);
  fr(c->lvl_ctl);
  fr(c->nodes);
  fr(c->ms);
  fr(c->perm);
  fr(c->rank);
  c->posm = nullptr;  // lives inside the nodes allocation
  fr(c->acc_t);
  fr(c->p2p_state);
  fr(c->wcounts);
  fr(c->tmp3);
  fr(c->gcost);
  fr(c->gorder);
  fr(c->wseed);
  c->wseed_n = ~0ull;
  c->gcost_n = 0;
}

// ---- peer-memory exchange set-up: cudaIpc handles of acc_t and of the flag array, all-gathered with NCCL
static void close_peers(Ctx* c) {
  for (int i = 0; i < 6 * P2P_MAX; ++i) {
    if (c->p2p_mapped[i]) cudaIpcCloseMemHandle(c->p2p_mapped[i]);
    c->p2p_mapped[i] = nullptr;
  }
  c->p2p_on = false;
  c->p2p_ready = false;
}

// all ranks of the communicator meet here (1-byte all-gather + stream synchronisation); errors are ignored — this is
// only ever a courtesy to the peers before memory they may still map is released
static void peer_barrier(Ctx* c) {
  if (c->world <= 1 || !c->nccl_comm || !g_nccl.all_gather) return;
  unsigned char* dev = nullptr;
  if (cudaMalloc(&dev, (size_t)c->world) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  if (g_nccl.all_gather(dev + c->rank_id, dev, 1, NCCL_UINT8, c->nccl_comm, c->stream) == 0) cudaStreamSynchronize(c->stream);
  cudaGetLastError();
  cudaFree(dev);
}

static int setup_peers(Ctx* c) {
  close_peers(c);
  c->p2p_ready = true;  // attempted for this allocation (success or not)
  static const bool disabled = getenv("KDNB_NO_P2P") != nullptr;
  const int W = c->world;
  // every rank must take the same decision and must reach both all-gathers whatever fails locally (a rank that
  // returned early would leave the others blocked in the collective): local failures vote ok = 0
  struct Rec {
    int ok;
    int pad[15];
    cudaIpcMemHandle_t acc, st, nodes, perm, ms, sexp;
  };
  static_assert(sizeof(Rec) == 448, "Rec");
  Rec mine;
  memset(&mine, 0, sizeof mine);
  mine.ok = (!disabled && W <= P2P_MAX) ? 1 : 0;
  if (cudaMemsetAsync(c->p2p_state, 0, P2P_WORDS * sizeof(uint32_t), c->stream) != cudaSuccess) mine.ok = 0;
  if (mine.ok && cudaIpcGetMemHandle(&mine.acc, c->acc_t) != cudaSuccess) mine.ok = 0;
  if (mine.ok && cudaIpcGetMemHandle(&mine.st, c->p2p_state) != cudaSuccess) mine.ok = 0;
  if (mine.ok && cudaIpcGetMemHandle(&mine.nodes, c->nodes) != cudaSuccess) mine.ok = 0;
  if (mine.ok && cudaIpcGetMemHandle(&mine.perm, c->perm) != cudaSuccess) mine.ok = 0;
  if (mine.ok && cudaIpcGetMemHandle(&mine.ms, c->ms) != cudaSuccess) mine.ok = 0;
  if (mine.ok && cudaIpcGetMemHandle(&mine.sexp, c->keys[1]) != cudaSuccess) mine.ok = 0;
  cudaGetLastError();
  Rec* dev = nullptr;
  int* dev2 = nullptr;
  std::string local_err;
  if (cudaMalloc(&dev, sizeof(Rec) * W) != cudaSuccess || cudaMalloc(&dev2, sizeof(int) * W) != cudaSuccess) {
    // without scratch this rank cannot even vote: the communicator is unusable for the set-up
    cudaGetLastError();
    if (dev) cudaFree(dev);
    return c->fail(KDNB_E_NOMEM, "peer set-up: cudaMalloc of the exchange scratch failed");
  }
  Rec all[KDNB_MAX_WORLD];  // W <= KDNB_MAX_WORLD (kdnb_comm_init)
  memset(all, 0, sizeof all);
  cudaError_t e = cudaMemcpyAsync(dev + c->rank_id, &mine, sizeof(Rec), cudaMemcpyHostToDevice, c->stream);
  int r = g_nccl.all_gather(dev + c->rank_id, dev, sizeof(Rec), NCCL_UINT8, c->nccl_comm, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(all, dev, sizeof(Rec) * W, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (r != 0) local_err = "ncclAllGather(ipc handles) failed";
  else if (e != cudaSuccess) local_err = std::string("ipc handle exchange: ") + cudaGetErrorString(e);
  bool ok = local_err.empty();
  for (int k = 0; ok && k < W; ++k) ok = ok && all[k].ok;
  P2P pp;
  memset(&pp, 0, sizeof pp);
  P2PBuild pb;
  memset(&pb, 0, sizeof pb);
  for (int k = 0; ok && k < W; ++k) {
    if (k == c->rank_id) {
      pp.acc[k] = c->acc_t;
      pp.flags[k] = c->p2p_state + 4;
      pb.nodes[k] = c->nodes, pb.perm[k] = c->perm, pb.ms[k] = c->ms, pb.state[k] = c->p2p_state;
      pb.sexp[k] = reinterpret_cast<uint32_t*>(c->keys[1]);
      continue;
    }
    void* m[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const cudaIpcMemHandle_t* h[6] = {&all[k].acc, &all[k].st, &all[k].nodes, &all[k].perm, &all[k].ms, &all[k].sexp};
    for (int q = 0; ok && q < 6; ++q) {
      if (cudaIpcOpenMemHandle(&m[q], *h[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        ok = false;
        cudaGetLastError();
      } else {
        c->p2p_mapped[6 * k + q] = m[q];  // (closed by close_peers, also after a partial failure)
      }
    }
    if (!ok) break;
    pp.acc[k] = reinterpret_cast<double*>(m[0]);
    pp.flags[k] = reinterpret_cast<uint32_t*>(m[1]) + 4;
    pb.nodes[k] = reinterpret_cast<WNode*>(m[2]);
    pb.perm[k] = reinterpret_cast<uint32_t*>(m[3]);
    pb.ms[k] = reinterpret_cast<double4*>(m[4]);
    pb.state[k] = reinterpret_cast<uint32_t*>(m[1]);
    pb.sexp[k] = reinterpret_cast<uint32_t*>(m[5]);
  }
  // second round: peer mode only if EVERY rank mapped every peer (otherwise all fall back to ncclAllGather)
  int okv = ok ? 1 : 0;
  int oks[KDNB_MAX_WORLD];
  memset(oks, 0, sizeof oks);
  e = cudaMemcpyAsync(dev2 + c->rank_id, &okv, sizeof(int), cudaMemcpyHostToDevice, c->stream);
  r = g_nccl.all_gather(dev2 + c->rank_id, dev2, sizeof(int), NCCL_UINT8, c->nccl_comm, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(oks, dev2, sizeof(int) * W, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(dev);
  cudaFree(dev2);
  if ((r != 0 || e != cudaSuccess) && local_err.empty()) local_err = "peer set-up vote failed";
  for (int k = 0; k < W; ++k) ok = ok && oks[k];
  if (!ok || !local_err.empty()) {
    const bool keep = c->p2p_ready;
    close_peers(c);
    c->p2p_ready = keep;
    if (!local_err.empty()) return c->fail(KDNB_E_NCCL, local_err);
    return 0;  // NCCL exchange
  }
  pp.epoch = c->p2p_state;
  pp.cta_done = c->p2p_state + 1;
  pp.stride = c->acc_stride;
  pp.world = W;
  pp.rank = c->rank_id;
  c->p2p = pp;
  pb.world = W;
  pb.rank = c->rank_id;
  c->p2pb = pb;
  c->p2p_on = true;
  return 0;
}

static void drop_graph_if_any(Ctx* c) {
  if (c->step_graph) cudaGraphExecDestroy(c->step_graph);
  c->step_graph = nullptr;
  c->graph_n = 0;
}

static uint64_t padded_slots(uint64_t n) { return n + 64ull * 64ull; }
// tree slots per rank: equal, warp-group aligned (64 = 32 lanes x 2 particles per lane) so that shards never split a warp
static uint64_t shard_slots_for(uint64_t n, int world) { return (((n + world - 1) / world) + 63) / 64 * 64; }

// (re)plan and (re)allocate for `n` particles
static int plan(Ctx* c, uint64_t n) {
  if (n > 0x7fffff00ull) return c->fail(KDNB_E_INVALID, "particle count exceeds the 32-bit index range of the device path");
  if (c->n != n) c->gcost_n = 0;  // (the recorded walk work stays a usable ordering hint for the same particle count)
  c->n = n;
  c->empty = false;
  c->n_nodes = subtree_nodes(n, c->mp, c->layout);
  c->ntiles = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE);
  int l0 = 0;
  while (((n + (1ull << l0) - 1) >> l0) > (uint64_t)BOT_CAP) ++l0;
  c->l0 = l0;
  c->tree_valid = c->acc_valid = c->map_valid = false;
  const uint64_t table = 2ull << l0;
  const uint64_t chunks = 3ull * (n / LVL_CHUNK + table + 8);
  const bool grow = n > c->cap || c->n_nodes > c->node_cap || table > c->table_cap || chunks > c->chunk_cap;
  if (grow) {
    drop_graph_if_any(c);
    free_all(c);
    c->cap = n;
    c->node_cap = c->n_nodes;
    c->table_cap = table;
    c->chunk_cap = chunks;
    int rc = 0;
    for (int d = 0; d < 3 && !rc; ++d) {
      rc = dev_alloc(c, &c->pos[d], n);
      if (!rc) rc = dev_alloc(c, &c->vel[d], n);
    }
    if (!rc) rc = dev_alloc(c, &c->mass, n);
    if (!rc) rc = dev_alloc(c, &c->radius, n);
    if (!rc) rc = dev_alloc(c, &c->aos, std::max<uint64_t>(n + 64, (c->n_nodes * sizeof(kdnb_node) + 63) / 64));  // +64: padded host shards
    if (!rc) rc = dev_alloc(c, &c->keys[0], 3 * n);
    if (!rc) rc = dev_alloc(c, &c->keys[1], 3 * n);
    if (!rc) rc = dev_alloc(c, &c->list[0], 3 * n);
    if (!rc) rc = dev_alloc(c, &c->list[1], 3 * n);
    if (!rc) rc = dev_alloc(c, &c->hist, 3ull * 256 * c->ntiles);
    if (!rc) rc = dev_alloc(c, &c->digit_tot, 3 * 256 + 16);
    if (!rc) c->flat = c->digit_tot + 3 * 256;
    if (!rc) c->dmask = c->flat + 4;
    if (!rc) rc = dev_alloc(c, &c->sort_state, SS_WORDS);
    if (!rc) rc = dev_alloc(c, &c->rk, 3 * n);
    if (!rc) rc = dev_alloc(c, &c->pm, n);
    if (!rc) rc = dev_alloc(c, &c->inv, n);
    if (!rc) rc = dev_alloc(c, &c->tseg, table);
    if (!rc) rc = dev_alloc(c, &c->lvl_status, chunks);
    if (!rc) rc = dev_alloc(c, &c->lvl_ctl, 72);
    auto cuda_rc = [&](cudaError_t e, const char* what) {
      return e == cudaSuccess ? 0 : c->fail(KDNB_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    };
    if (!rc) rc = cuda_rc(cudaMemsetAsync(c->lvl_status, 0, chunks * sizeof(uint64_t), c->stream), "cudaMemsetAsync(lvl_status)");
    if (!rc) rc = cuda_rc(cudaMemsetAsync(c->lvl_ctl, 0, 72 * sizeof(uint32_t), c->stream), "cudaMemsetAsync(lvl_ctl)");
    if (!rc) rc = dev_alloc(c, &c->nodes, c->n_nodes + (n + 1) / 2 + 1);  // node records, then the tree-ordered particles (one pool: walk.cu)
    if (!rc) rc = dev_alloc(c, &c->ms, c->n_nodes);
    if (!rc) rc = dev_alloc(c, &c->perm, n);
    if (!rc) rc = dev_alloc(c, &c->rank, n);
    if (!rc) c->posm = reinterpret_cast<PosM*>(c->nodes + c->n_nodes);
    if (!rc) rc = dev_alloc(c, &c->acc_t, 2 * 3 * padded_slots(n));  // two buffers: peer exchange alternates by step parity
    if (!rc) rc = dev_alloc(c, &c->p2p_state, P2P_WORDS);
    if (!rc && (c->flags & KDNB_FLAG_WALK_COUNTS)) rc = dev_alloc(c, &c->wcounts, 4 * n);
    if (!rc) rc = dev_alloc(c, &c->tmp3, 4 * n);
    if (!rc) rc = dev_alloc(c, &c->gcost, n / 32 + 2);
    if (!rc) rc = dev_alloc(c, &c->gorder, n / 32 + 2);
    if (!rc) rc = dev_alloc(c, &c->wseed, 64);
    if (rc) {
      const std::string why = c->err;
      free_all(c);
      c->cap = c->node_cap = c->table_cap = c->chunk_cap = 0;
      c->n = c->planned_n = 0;
      c->err = why;
      return rc;
    }
  }
  c->posm = reinterpret_cast<PosM*>(c->nodes + c->n_nodes);
  if (grow || c->planned_n != n) {
    KDNB_CUDA_TRY(c, cudaMemsetAsync(c->acc_t, 0, 2 * 3 * padded_slots(n) * sizeof(double), c->stream));
    init_unused_nodes(c);  // slots the build never writes keep the reference's default Leaf{0, NEGS}
    c->planned_n = n;
  }
  c->acc_stride = 3 * padded_slots(c->cap);
  if (grow || c->wseed_n != n) {
    // The top of the tree is the same descent for every group of the walk: node indices of depths 0-5 in heap order
    // (closed form: the left child follows its parent, the right child follows the left subtree), usable when every
    // node of depths 0-4 is internal.  Identical on every rank, whatever part of the tree it builds.
    uint32_t node[64];
    uint64_t len[64];
    memset(node, 0, sizeof node);
    memset(len, 0, sizeof len);
    node[1] = 0, len[1] = n;
    bool ok = true;
    for (int h = 1; h < 32 && ok; ++h) {
      if (len[h] <= c->mp) ok = false;
      const uint64_t half = len[h] / 2;
      node[2 * h] = node[h] + 1, len[2 * h] = half;
      node[2 * h + 1] = node[h] + 1 + (uint32_t)subtree_nodes(half, c->mp, c->layout), len[2 * h + 1] = len[h] - half;
    }
    static const bool off = [] { const char* s = getenv("KDNB_WALK_SEED"); return s && atoi(s) == 0; }();
    c->wseed_ok = ok && !off;
    if (c->wseed_ok) {
      KDNB_CUDA_TRY(c, cudaMemcpyAsync(c->wseed, node + 1, 63 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
      KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));  // (node[] is a stack array)
    }
    c->wseed_n = n;
  }
  if (c->world > 1) {
    c->shard_slots = shard_slots_for(n, c->world);
    if (grow || !c->p2p_ready) {
      if (int rc = setup_peers(c)) return rc;
    }
  }
  return 0;
}

static int exchange(Ctx* c) {
  if (c->world <= 1) return 0;
  if (c->p2p_on) return p2p_wait_step(c);  // the walk kernel already stored into every peer: wait for all flags
  const size_t count = (size_t)c->shard_slots * 3;
  int r = g_nccl.all_gather(c->acc_t + (size_t)c->rank_id * count, c->acc_t, count, NCCL_FLOAT64, c->nccl_comm, c->stream);
  if (r != 0)
    return c->fail(KDNB_E_NCCL, std::string("ncclAllGather: ") + (g_nccl.get_error_string ? g_nccl.get_error_string(r) : "error"));
  return 0;
}

// multi-GPU, peer mode: the wait kernel gives up after ~15 s when a peer never publishes its accelerations (walk.cu)
// and records it; the step then ran on incomplete accelerations, so the caller must hear about it
static int check_peer_timeout(Ctx* c) {
  if (c->world <= 1 || !c->p2p_on || !c->p2p_state) return 0;
  uint32_t err = 0;
  KDNB_CUDA_TRY(c, cudaMemcpyAsync(&err, c->p2p_state + 2, sizeof(err), cudaMemcpyDeviceToHost, c->stream));
  KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (err) return c->fail(KDNB_E_CUDA, "multi-GPU exchange: a peer did not publish its accelerations in time (results are incomplete)");
  return 0;
}

// One step.  KDNB_FLAG_PROFILE: the four stage boundaries are recorded as events — external event-record nodes when the
// step is being captured into the step graph, so that the stage times come from the SAME execution mode (graph replay)
// as the timed region of bench.py; stage_collect() reads them after every profiled step.
static int one_step(Ctx* c, double dt, bool capturing = false) {
  const bool prof = (c->flags & KDNB_FLAG_PROFILE) != 0;
  if (prof && !c->pev[0]) {
    for (int k = 0; k < 5; ++k) KDNB_CUDA_TRY(c, cudaEventCreate(&c->pev[k]));
  }
  auto mark = [&](int k) {
    if (prof) cudaEventRecordWithFlags(c->pev[k], c->stream, capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
  };
  mark(0);
  walk_order_fork(c);                     // (the walk's launch order, from the previous walk: on a side stream next to the build)
  if (int rc = build_tree(c)) return rc;  // indices reset + build_tree_par4 (:641-643)
  mark(1);
  if (int rc = walk(c)) return rc;        // calc_accel for every particle (:647)
  mark(2);
  if (int rc = exchange(c)) return rc;
  mark(3);
  if (int rc = kick_drift(c, dt)) return rc;  // (:649-662)
  mark(4);
  return 0;
}

// profiled contexts: wait for the step just enqueued and add its stage times to the sums
static int stage_collect(Ctx* c) {
  if (!(c->flags & KDNB_FLAG_PROFILE) || !c->pev[0]) return 0;
  KDNB_CUDA_TRY(c, cudaEventSynchronize(c->pev[4]));
  for (int k = 0; k < 4; ++k) {
    float t = 0.f;
    KDNB_CUDA_TRY(c, cudaEventElapsedTime(&t, c->pev[k], c->pev[k + 1]));
    c->stage_acc[k] += t;
  }
  c->ev_steps++;
  return 0;
}

// ------------------------------------------------------------------------------------------------ C ABI

extern "C" {

int kdnb_version(void) { return KDNB_VERSION; }

const char* kdnb_last_error(const kdnb_ctx* ctx) { return ctx ? ctx->c.err.c_str() : g_create_error.c_str(); }

kdnb_ctx* kdnb_create(const kdnb_config* cfg) {
  kdnb_config k;
  memset(&k, 0, sizeof k);
  if (cfg) memcpy(&k, cfg, std::min<size_t>(sizeof k, cfg->struct_size ? cfg->struct_size : sizeof k));
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                     " (libkdnb has no CPU fallback)";
    return nullptr;
  }
  if (k.device < 0 || k.device >= ndev) {
    g_create_error = "device ordinal out of range";
    return nullptr;
  }
  const uint32_t mp = k.max_parts ? k.max_parts : 8;
  if (mp < 4 || mp > 32) {
    g_create_error = "max_parts must be in 4..32";
    return nullptr;
  }
  if (k.layout != KDNB_LAYOUT_PADDED && k.layout != KDNB_LAYOUT_DENSE) {
    g_create_error = "unknown layout";
    return nullptr;
  }
  kdnb_ctx* h = new (std::nothrow) kdnb_ctx();
  if (!h) {
    g_create_error = "out of host memory";
    return nullptr;
  }
  Ctx* c = &h->c;
  c->device = k.device;
  c->mp = mp;
  c->layout = k.layout;
  c->theta = k.theta != 0.0 ? k.theta : 0.3;
  c->theta2 = c->theta * c->theta;  // THETA * THETA (array_kd_tree.rs:606)
  c->flags = k.flags;
  if ((e = cudaSetDevice(c->device)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&c->order_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreate(&c->sw_begin)) != cudaSuccess || (e = cudaEventCreate(&c->sw_end)) != cudaSuccess) {
    g_create_error = std::string("CUDA init: ") + cudaGetErrorString(e);
    delete h;
    return nullptr;
  }
  cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device);
  c->use_pdl = getenv("KDNB_PDL") != nullptr;  // measured: no gain on top of CUDA-graph replay (profiles/README.md)
  return h;
}

void kdnb_destroy(kdnb_ctx* ctx) {
  if (!ctx) return;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  drop_graph_if_any(c);
  free_all(c);  // (world > 1 in peer mode: meets the other ranks before the exported buffers go — destroy is collective)
  if (c->nccl_comm && g_nccl.comm_destroy) g_nccl.comm_destroy(c->nccl_comm);
  c->nccl_comm = nullptr;
  if (c->l2_scratch) cudaFree(c->l2_scratch);
  for (cudaEvent_t x : c->pev)
    if (x) cudaEventDestroy(x);
  if (c->sw_begin) cudaEventDestroy(c->sw_begin);
  if (c->sw_end) cudaEventDestroy(c->sw_end);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->side_stream) cudaStreamDestroy(c->side_stream);
  if (c->order_stream) cudaStreamDestroy(c->order_stream);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  delete ctx;
}

#define CTX_OR_FAIL(ctx)                 \
  if (!(ctx)) return KDNB_E_INVALID;     \
  Ctx* c = &(ctx)->c;                    \
  KDNB_CUDA_TRY(c, cudaSetDevice(c->device))

#define NEED_PARTICLES(c) \
  if ((c)->n == 0 && !(c)->empty) return (c)->fail(KDNB_E_INVALID, "no particles uploaded")

// `bodies` is empty: the reference's simple_sim then owns an empty acc / indices and the one-node tree of
// allocate_node_vec(0) (array_kd_tree.rs:45-60), whose node 0 the build writes as Leaf{0, [0; MAX_PARTS]} (:524-529),
// and every stage loops over nothing.  No device work.
static int upload_empty(Ctx* c) {
  c->n = 0;
  c->empty = true;
  c->n_nodes = 1;
  c->planned_n = 0;
  c->tree_valid = c->acc_valid = c->map_valid = false;
  return 0;
}

int kdnb_upload_particles(kdnb_ctx* ctx, const kdnb_particle* aos, uint64_t count) {
  CTX_OR_FAIL(ctx);
  if (count == 0) return upload_empty(c);
  if (!aos) return c->fail(KDNB_E_INVALID, "null particle array");
  if (int rc = plan(c, count)) return rc;
  KDNB_CUDA_TRY(c, cudaMemcpyAsync(c->aos, aos, count * sizeof(kdnb_particle), cudaMemcpyHostToDevice, c->stream));
  return aos_to_soa(c);
}

int kdnb_download_particles(kdnb_ctx* ctx, kdnb_particle* out, uint64_t capacity) {
  CTX_OR_FAIL(ctx);
  NEED_PARTICLES(c);
  if (c->empty) return 0;
  if (!out) return c->fail(KDNB_E_INVALID, "null output array");
  if (capacity < c->n) return c->fail(KDNB_E_CAPACITY, "output array smaller than the particle count");
  if (int rc = soa_to_aos(c)) return rc;
  KDNB_CUDA_TRY(c, cudaMemcpyAsync(out, c->aos, c->n * sizeof(kdnb_particle), cudaMemcpyDeviceToHost, c->stream));
  KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return check_peer_timeout(c);
}

uint64_t kdnb_particle_count(const kdnb_ctx* ctx) { return ctx ? ctx->c.n : 0; }

int kdnb_build_tree(kdnb_ctx* ctx) {
  CTX_OR_FAIL(ctx);
  NEED_PARTICLES(c);
  if (c->empty) {
    c->tree_valid = c->map_valid = true;
    return 0;
  }
  return build_tree(c);
}

int kdnb_calc_accel(kdnb_ctx* ctx) {
  CTX_OR_FAIL(ctx);
  if (!c->tree_valid) return c->fail(KDNB_E_INVALID, "kdnb_calc_accel needs kdnb_build_tree on the current positions first");
  if (c->empty) {
    c->acc_valid = true;
    return 0;
  }
  if (int rc = walk(c)) return rc;
  return exchange(c);
}

int kdnb_kick_drift(kdnb_ctx* ctx, double dt) {
  CTX_OR_FAIL(ctx);
  NEED_PARTICLES(c);
  if (!c->map_valid) return c->fail(KDNB_E_INVALID, "kdnb_kick_drift needs the particle->slot map of a previous kdnb_build_tree");
  if (c->empty) {
    c->tree_valid = false;
    return 0;
  }
  return kick_drift(c, dt);
}

int kdnb_simple_sim(kdnb_ctx* ctx, double dt, int64_t steps) {
  CTX_OR_FAIL(ctx);
  NEED_PARTICLES(c);
  if (c->empty || steps <= 0) return 0;  // `for _ in 0..steps` (array_kd_tree.rs:632) runs nothing
  int64_t s = 0;
  // Launch-bound regime (N <= ~1M: ~35 kernels of 5-35 us per step): replay the step as one CUDA graph.  A graph is
  // captured by a call of >= 3 steps, or by the third consecutive call with the same (n, dt) — a caller stepping one
  // step per call, like the reference's own loop around its closure; once it exists every step of a matching call
  // replays it.  The step before a capture always runs as plain launches (lazy one-time setup: function attributes,
  // NCCL connections).  Profiled contexts replay the same graph (with event-record nodes at the stage boundaries) and
  // wait for every step to read them.
  static const bool no_graph = getenv("KDNB_NO_GRAPH") != nullptr;
  const bool prof = (c->flags & KDNB_FLAG_PROFILE) != 0;
  const bool can_graph = !no_graph;
  const bool have = c->step_graph && c->graph_n == c->n && c->graph_dt == dt && c->graph_world == c->world;
  if (c->seen_n == c->n && c->seen_dt == dt) {
    c->seen_calls++;
  } else {
    c->seen_n = c->n;
    c->seen_dt = dt;
    c->seen_calls = 1;
  }
  const bool use_graph = can_graph && (have || steps >= 3 || c->seen_calls >= 3);
  if (use_graph) {
    if (!have) {
      if (int rc = one_step(c, dt)) return rc;
      if (int rc = stage_collect(c)) return rc;
      s = 1;
      drop_graph_if_any(c);
      const uint64_t l0 = c->launches;
      KDNB_CUDA_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      const int rc = one_step(c, dt, true);
      cudaGraph_t g = nullptr;
      cudaError_t e = cudaStreamEndCapture(c->stream, &g);
      c->graph_launches = c->launches - l0;
      c->launches = l0;  // the capture launched nothing
      if (rc || e != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        if (rc) return rc;
        return c->fail(KDNB_E_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
      }
      e = cudaGraphInstantiate(&c->step_graph, g, 0);
      cudaGraphDestroy(g);
      if (e != cudaSuccess) {
        c->step_graph = nullptr;
        return c->fail(KDNB_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
      }
      c->graph_n = c->n;
      c->graph_dt = dt;
      c->graph_world = c->world;
    }
    for (; s < steps; ++s) {
      KDNB_CUDA_TRY(c, cudaGraphLaunch(c->step_graph, c->stream));
      c->launches += c->graph_launches;
      if (prof)
        if (int rc = stage_collect(c)) return rc;
    }
    return 0;
  }
  for (; s < steps; ++s) {
    if (int rc = one_step(c, dt)) return rc;
    if (prof)
      if (int rc = stage_collect(c)) return rc;
  }
  return 0;
}

// ---- sharded host buffers (multi-GPU): rank r holds particles [first, first+shard) of the global array
static uint64_t host_shard(uint64_t total, int world) { return (total + world - 1) / world; }

int kdnb_host_shard_range(uint64_t total, int rank, int world_size, uint64_t* first, uint64_t* count) {
  if (world_size < 1 || rank < 0 || rank >= world_size || !first || !count) return KDNB_E_INVALID;
  const uint64_t s = host_shard(total, world_size);
  *first = std::min<uint64_t>(total, (uint64_t)rank * s);
  *count = std::min<uint64_t>(total, (uint64_t)(rank + 1) * s) - *first;
  return 0;
}

int kdnb_upload_particles_sharded(kdnb_ctx* ctx, const kdnb_particle* shard, uint64_t total) {
  CTX_OR_FAIL(ctx);
  if (c->world <= 1) return kdnb_upload_particles(ctx, shard, total);
  if (total == 0) return upload_empty(c);
  if (!shard) return c->fail(KDNB_E_INVALID, "null particle array");
  if (int rc = plan(c, total)) return rc;
  const uint64_t s = host_shard(total, c->world);
  uint64_t first = 0, cnt = 0;
  kdnb_host_shard_range(total, c->rank_id, c->world, &first, &cnt);
  // own shard over PCIe, everybody else's over NVLink (in-place ncclAllGather of equal, padded shards)
  if (cnt) KDNB_CUDA_TRY(c, cudaMemcpyAsync(c->aos + (size_t)c->rank_id * s, shard, cnt * sizeof(kdnb_particle), cudaMemcpyHostToDevice, c->stream));
  int r = g_nccl.all_gather(c->aos + (size_t)c->rank_id * s, c->aos, s * sizeof(kdnb_particle), NCCL_UINT8, c->nccl_comm, c->stream);
  if (r != 0) return c->fail(KDNB_E_NCCL, "ncclAllGather(particles) failed");
  return aos_to_soa(c);
}

int kdnb_download_particles_sharded(kdnb_ctx* ctx, kdnb_particle* shard_out) {
  CTX_OR_FAIL(ctx);
  NEED_PARTICLES(c);
  if (c->empty) return 0;
  if (!shard_out) return c->fail(KDNB_E_INVALID, "null output array");
  uint64_t first = 0, cnt = 0;
  kdnb_host_shard_range(c->n, c->rank_id, c->world, &first, &cnt);
  if (int rc = soa_to_aos(c)) return rc;  // every replica holds the full, bit-identical state
  if (cnt) KDNB_CUDA_TRY(c, cudaMemcpyAsync(shard_out, c->aos + first, cnt * sizeof(kdnb_particle), cudaMemcpyDeviceToHost, c->stream));
  KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return check_peer_timeout(c);
}

int kdnb_simple_sim_bodies_sharded(kdnb_ctx* ctx, kdnb_particle* shard, uint64_t total, double dt, int64_t steps) {
  if (int rc = kdnb_upload_particles_sharded(ctx, shard, total)) return rc;
  if (int rc = kdnb_simple_sim(ctx, dt, steps)) return rc;
  return kdnb_download_particles_sharded(ctx, shard);
}

int kdnb_simple_sim_bodies(kdnb_ctx* ctx, kdnb_particle* bodies, uint64_t count, double dt, int64_t steps) {
  if (int rc = kdnb_upload_particles(ctx, bodies, count)) return rc;
  if (int rc = kdnb_simple_sim(ctx, dt, steps)) return rc;
  return kdnb_download_particles(ctx, bodies, count);
}

// ---- the Sequential crate's SIMD particle surface (simd_particle.rs:3-8, simd_kd_tree.rs:169-202)
static int simd_staging(Ctx* c, uint64_t count) {
  if (c->aos_simd && c->aos_simd_cap >= count) return 0;
  if (int rc = dev_alloc(c, &c->aos_simd, count)) return rc;
  c->aos_simd_cap = count;
  return 0;
}

// a record whose padding lane was not 0 was uploaded (flag raised by simd_to_soa_kernel)
static int check_simd_lanes(Ctx* c) {
  if (!c->lvl_ctl) return 0;
  uint32_t bad = 0;
  KDNB_CUDA_TRY(c, cudaMemcpyAsync(&bad, c->lvl_ctl + 66, sizeof(bad), cudaMemcpyDeviceToHost, c->stream));
  KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (bad) {
    KDNB_CUDA_TRY(c, cudaMemsetAsync(c->lvl_ctl + 66, 0, sizeof(uint32_t), c->stream));
    return c->fail(KDNB_E_INVALID, "kdnb_particle_simd: lane 3 of p / v must be 0 (it is padding in the reference's SIMD particle)");
  }
  return 0;
}

int kdnb_upload_particles_simd(kdnb_ctx* ctx, const kdnb_particle_simd* aos, uint64_t count) {
  CTX_OR_FAIL(ctx);
  if (count == 0) return upload_empty(c);
  if (!aos) return c->fail(KDNB_E_INVALID, "null particle array");
  if (int rc = plan(c, count)) return rc;
  if (int rc = simd_staging(c, count)) return rc;
  KDNB_CUDA_TRY(c, cudaMemcpyAsync(c->aos_simd, aos, count * sizeof(kdnb_particle_simd), cudaMemcpyHostToDevice, c->stream));
  return simd_to_soa(c, c->aos_simd);
}

int kdnb_download_particles_simd(kdnb_ctx* ctx, kdnb_particle_simd* out, uint64_t capacity) {
  CTX_OR_FAIL(ctx);
  NEED_PARTICLES(c);
  if (c->empty) return 0;
  if (!out) return c->fail(KDNB_E_INVALID, "null output array");
  if (capacity < c->n) return c->fail(KDNB_E_CAPACITY, "output array smaller than the particle count");
  if (int rc = simd_staging(c, c->n)) return rc;
  if (int rc = soa_to_simd(c, c->aos_simd)) return rc;
  KDNB_CUDA_TRY(c, cudaMemcpyAsync(out, c->aos_simd, c->n * sizeof(kdnb_particle_simd), cudaMemcpyDeviceToHost, c->stream));
  KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (int rc = check_simd_lanes(c)) return rc;
  return check_peer_timeout(c);
}

int kdnb_simple_sim_bodies_simd(kdnb_ctx* ctx, kdnb_particle_simd* bodies, uint64_t count, double dt, int64_t steps) {
  if (int rc = kdnb_upload_particles_simd(ctx, bodies, count)) return rc;
  if (int rc = check_simd_lanes(&ctx->c)) return rc;  // before the caller's records are overwritten
  if (int rc = kdnb_simple_sim(ctx, dt, steps)) return rc;
  return kdnb_download_particles_simd(ctx, bodies, count);
}

int kdnb_simple_sim_host(const kdnb_config* cfg, kdnb_particle* bodies, uint64_t count, double dt, int64_t steps) {
  kdnb_ctx* h = kdnb_create(cfg);
  if (!h) return KDNB_E_CUDA;
  int rc = kdnb_simple_sim_bodies(h, bodies, count, dt, steps);
  if (rc) g_create_error = h->c.err;
  kdnb_destroy(h);
  return rc;
}

int kdnb_synchronize(kdnb_ctx* ctx) {
  CTX_OR_FAIL(ctx);
  KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (int rc = check_peer_timeout(c)) return rc;
  if (c->lvl_ctl) {  // a look-back of the level partitions that gave up instead of spinning for ever (build.cu)
    uint32_t err = 0;
    KDNB_CUDA_TRY(c, cudaMemcpyAsync(&err, c->lvl_ctl + 65, sizeof(err), cudaMemcpyDeviceToHost, c->stream));
    KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (err) return c->fail(KDNB_E_CUDA, "tree build: a level partition timed out waiting for a predecessor chunk");
  }
  return 0;
}

int kdnb_download_accel(kdnb_ctx* ctx, double* acc) {
  CTX_OR_FAIL(ctx);
  NEED_PARTICLES(c);
  if (c->empty) return 0;
  if (!acc) return c->fail(KDNB_E_INVALID, "null output array");
  if (!c->tree_valid || !c->acc_valid) {
    // before the first calc_accel and after every kick the reference's acc vector is all zeros (:624-627, :659-661);
    // the device keeps that state as a flag (acc_valid), not as 24 bytes per particle of zeros
    memset(acc, 0, 3 * c->n * sizeof(double));
    return 0;
  }
  if (int rc = gather_acc(c, c->tmp3)) return rc;
  KDNB_CUDA_TRY(c, cudaMemcpyAsync(acc, c->tmp3, 3 * c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int kdnb_upload_accel(kdnb_ctx* ctx, const double* acc) {
  CTX_OR_FAIL(ctx);
  if (!c->tree_valid) return c->fail(KDNB_E_INVALID, "kdnb_upload_accel needs a built tree (particle->slot map)");
  if (c->empty) {
    c->acc_valid = true;
    return 0;
  }
  if (!acc) return c->fail(KDNB_E_INVALID, "null input array");
  KDNB_CUDA_TRY(c, cudaMemcpyAsync(c->tmp3, acc, 3 * c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (int rc = scatter_acc(c, c->tmp3)) return rc;
  c->acc_valid = true;
  return 0;
}

int kdnb_download_tree(kdnb_ctx* ctx, kdnb_node* nodes, uint64_t cap, uint64_t* n_nodes, uint64_t* indices) {
  CTX_OR_FAIL(ctx);
  if (!c->tree_valid) return c->fail(KDNB_E_INVALID, "no tree built for the current positions");
  if (n_nodes) *n_nodes = c->n_nodes;
  if (c->empty) {  // allocate_node_vec(0) is one node; the build wrote it as an empty leaf (array_kd_tree.rs:524-529)
    if (nodes) {
      if (cap < 1) return c->fail(KDNB_E_CAPACITY, "node array smaller than kdnb_node_count()");
      memset(&nodes[0], 0, sizeof(kdnb_node));
      nodes[0].kind = KDNB_LEAF;
    }
    return 0;
  }
  if (nodes) {
    if (cap < c->n_nodes) return c->fail(KDNB_E_CAPACITY, "node array smaller than kdnb_node_count()");
    kdnb_node* dev = reinterpret_cast<kdnb_node*>(c->aos);  // staging buffer is sized for this in plan()
    if (int rc = export_tree(c, dev)) return rc;
    KDNB_CUDA_TRY(c, cudaMemcpyAsync(nodes, dev, c->n_nodes * sizeof(kdnb_node), cudaMemcpyDeviceToHost, c->stream));
  }
  KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (indices) {
    // the device holds u32 slots; land them in the upper half of the caller's u64 array and widen front to back in
    // place (entry i is written over bytes [8i, 8i+8), which lie at or below its own packed source at byte 4n + 4i
    // and strictly below every packed word still to be read), so no host staging buffer is needed at any N
    unsigned char* bytes = reinterpret_cast<unsigned char*>(indices);
    KDNB_CUDA_TRY(c, cudaMemcpy(bytes + 4 * c->n, c->perm, c->n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < c->n; ++i) {
      uint32_t v;
      memcpy(&v, bytes + 4 * c->n + 4 * i, sizeof v);
      const uint64_t w = v;
      memcpy(bytes + 8 * i, &w, sizeof w);
    }
  }
  return 0;
}

int kdnb_download_walk_counts(kdnb_ctx* ctx, uint64_t* counts) {
  CTX_OR_FAIL(ctx);
  if (!(c->flags & KDNB_FLAG_WALK_COUNTS)) return c->fail(KDNB_E_INVALID, "context was created without KDNB_FLAG_WALK_COUNTS");
  if (!c->acc_valid || !c->tree_valid) return c->fail(KDNB_E_INVALID, "no walk results for the current tree");
  if (c->empty) return 0;
  unsigned long long* dev = reinterpret_cast<unsigned long long*>(c->tmp3);
  if (int rc = gather_counts(c, dev)) return rc;
  KDNB_CUDA_TRY(c, cudaMemcpyAsync(counts, dev, 4 * c->n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
  KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return 0;
}

uint64_t kdnb_nodes_needed(uint64_t num_parts, uint32_t max_parts) {
  if (max_parts < 2) return 0;
  return nodes_needed_padded(num_parts, max_parts);
}

uint64_t kdnb_node_count(const kdnb_ctx* ctx) { return ctx ? ctx->c.n_nodes : 0; }

int kdnb_shard_range(uint64_t count, int rank, int world_size, uint64_t* begin, uint64_t* end) {
  if (world_size < 1 || rank < 0 || rank >= world_size || !begin || !end) return KDNB_E_INVALID;
  const uint64_t shard = shard_slots_for(count, world_size);
  *begin = std::min<uint64_t>(count, (uint64_t)rank * shard);
  *end = std::min<uint64_t>(count, (uint64_t)(rank + 1) * shard);
  return 0;
}

int kdnb_build_shard_plan(uint64_t count, uint32_t max_parts, int layout, int rank, int world_size, uint64_t* first_slot,
                          uint64_t* slots, uint64_t* first_node, uint64_t* nodes) {
  if (world_size < 1 || rank < 0 || rank >= world_size || max_parts < 4 || max_parts > 32 || !first_slot || !slots || !first_node || !nodes)
    return KDNB_E_INVALID;
  if (layout != KDNB_LAYOUT_PADDED && layout != KDNB_LAYOUT_DENSE) return KDNB_E_INVALID;
  int k = 0;
  while ((1 << k) < world_size) ++k;
  if ((1 << k) != world_size) return KDNB_E_INVALID;
  uint64_t a = 0, len = 0, node = 0;
  subtree_of(count, max_parts, layout, k, (uint32_t)rank, &a, &len, &node);
  *first_slot = a, *slots = len, *first_node = node, *nodes = subtree_nodes(len, max_parts, layout);
  return 0;
}

int kdnb_comm_unique_id(void* id_out) {
  std::string why;
  if (!id_out || !load_nccl(&why)) {
    g_create_error = why;
    return KDNB_E_NCCL;
  }
  NcclId id;
  int r = g_nccl.get_unique_id(&id);
  if (r != 0) return KDNB_E_NCCL;
  memcpy(id_out, &id, sizeof id);
  return 0;
}

int kdnb_comm_init(kdnb_ctx* ctx, const void* id_bytes, int rank, int world_size) {
  CTX_OR_FAIL(ctx);
  if (world_size < 1 || rank < 0 || rank >= world_size || world_size > KDNB_MAX_WORLD) return c->fail(KDNB_E_INVALID, "bad rank / world size");
  if (c->nccl_comm) {  // joining again (another communicator or world size): leave the previous one first
    KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    drop_graph_if_any(c);
    close_peers(c);
    if (g_nccl.comm_destroy) g_nccl.comm_destroy(c->nccl_comm);
    c->nccl_comm = nullptr;
  }
  if (world_size == 1) {
    c->rank_id = 0;
    c->world = 1;
    return 0;
  }
  std::string why;
  if (!id_bytes || !load_nccl(&why)) return c->fail(KDNB_E_NCCL, why.empty() ? "null NCCL id" : why);
  NcclId id;
  memcpy(&id, id_bytes, sizeof id);
  void* comm = nullptr;
  int r = g_nccl.comm_init_rank(&comm, world_size, id, rank);
  if (r != 0) {
    c->rank_id = 0;  // (a failed re-join leaves a working single-rank context, not world > 1 without a communicator)
    c->world = 1;
    return c->fail(KDNB_E_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.get_error_string ? g_nccl.get_error_string(r) : "error"));
  }
  c->nccl_comm = comm;
  c->rank_id = rank;
  c->world = world_size;
  if (c->n) {
    c->shard_slots = shard_slots_for(c->n, c->world);
    c->acc_stride = 3 * padded_slots(c->cap);
    if (int rc2 = setup_peers(c)) return rc2;
  }
  return 0;
}

int kdnb_stage_ms(kdnb_ctx* ctx, double ms_out[KDNB_STAGE_COUNT], uint64_t* steps_out) {
  CTX_OR_FAIL(ctx);
  if (!(c->flags & KDNB_FLAG_PROFILE)) return c->fail(KDNB_E_INVALID, "context was created without KDNB_FLAG_PROFILE");
  KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < KDNB_STAGE_COUNT; ++k) ms_out[k] = 0.0;
  ms_out[KDNB_STAGE_BUILD] = c->stage_acc[0];
  ms_out[KDNB_STAGE_WALK] = c->stage_acc[1];
  ms_out[KDNB_STAGE_EXCHANGE] = c->stage_acc[2];
  ms_out[KDNB_STAGE_KICK] = c->stage_acc[3];
  if (steps_out) *steps_out = c->ev_steps;
  return 0;
}

int kdnb_stage_reset(kdnb_ctx* ctx) {
  CTX_OR_FAIL(ctx);
  KDNB_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->ev_steps = 0;
  for (double& x : c->stage_acc) x = 0.0;
  return 0;
}

uint64_t kdnb_launch_count(const kdnb_ctx* ctx) { return ctx ? ctx->c.launches : 0; }

int kdnb_quickstat_index(kdnb_ctx* ctx, const double* vals, uint64_t n_vals, uint64_t* indices, uint64_t count,
                         uint64_t goal, double* device_ms) {
  CTX_OR_FAIL(ctx);
  if (!vals || !indices) return c->fail(KDNB_E_INVALID, "null array");
  if (count == 0 || goal >= count) return c->fail(KDNB_E_INVALID, "goal out of range");  // the reference indexes out of bounds
  if (n_vals == 0 || n_vals > 0xffffffffull || count > 0xffffff00ull)
    return c->fail(KDNB_E_INVALID, "value count exceeds the 32-bit index range of the device path");
  for (uint64_t i = 0; i < count; ++i)
    if (indices[i] >= n_vals) return c->fail(KDNB_E_INVALID, "index out of range of vals");
  KDNB_CUDA_TRY(c, cudaSetDevice(c->device));
  const uint64_t ntiles = (count + 2047) / 2048;
  double* d_vals = nullptr;
  uint64_t *d_idx = nullptr, *d_keys = nullptr, *d_state = nullptr;
  uint32_t* d_idx32 = nullptr;
  uint2* d_tiles = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_vals), cudaFree(d_idx), cudaFree(d_keys), cudaFree(d_state), cudaFree(d_idx32), cudaFree(d_tiles);
  };
  cudaError_t e = cudaSuccess;
  if ((e = cudaMalloc(&d_vals, n_vals * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(&d_idx, count * sizeof(uint64_t))) != cudaSuccess ||
      (e = cudaMalloc(&d_keys, count * sizeof(uint64_t))) != cudaSuccess ||
      (e = cudaMalloc(&d_state, 8 * sizeof(uint64_t) + 256 * sizeof(uint32_t))) != cudaSuccess ||
      (e = cudaMalloc(&d_idx32, count * sizeof(uint32_t))) != cudaSuccess ||
      (e = cudaMalloc(&d_tiles, ntiles * sizeof(uint2))) != cudaSuccess) {
    cleanup();
    return c->fail(KDNB_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  int rc = 0;
  float ms = 0.f;
  if ((e = cudaMemcpyAsync(d_vals, vals, n_vals * sizeof(double), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(d_idx, indices, count * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) {
    rc = c->fail(KDNB_E_CUDA, std::string("upload: ") + cudaGetErrorString(e));
  }
  if (!rc) {
    cudaEventRecord(c->sw_begin, c->stream);
    // (the u64 index buffer is dead once sel_keys has narrowed it to u32: it receives the result)
    rc = select_run(c, d_vals, d_idx, count, goal, d_keys, d_idx32, d_tiles, d_state, d_idx);
    cudaEventRecord(c->sw_end, c->stream);
  }
  if (!rc && (e = cudaMemcpyAsync(indices, d_idx, count * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess)
    rc = c->fail(KDNB_E_CUDA, std::string("download: ") + cudaGetErrorString(e));
  if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess && !rc)
    rc = c->fail(KDNB_E_CUDA, std::string("select: ") + cudaGetErrorString(e));
  if (!rc && device_ms && cudaEventElapsedTime(&ms, c->sw_begin, c->sw_end) == cudaSuccess) *device_ms = ms;
  cleanup();
  return rc;
}

int kdnb_measure_fp64_peak(kdnb_ctx* ctx, double* tflops_out) {
  CTX_OR_FAIL(ctx);
  if (!tflops_out) return c->fail(KDNB_E_INVALID, "null output");
  return measure_fp64_peak(c, tflops_out);
}

int kdnb_flush_l2(kdnb_ctx* ctx) {
  CTX_OR_FAIL(ctx);
  const size_t bytes = 256ull << 20;
  if (!c->l2_scratch) {
    cudaError_t e = cudaMalloc(&c->l2_scratch, bytes);
    if (e != cudaSuccess) return c->fail(KDNB_E_NOMEM, "cudaMalloc(l2 scratch)");
  }
  KDNB_CUDA_TRY(c, cudaMemsetAsync(c->l2_scratch, (int)(c->launches & 0xff), bytes, c->stream));
  return 0;
}

int kdnb_device_ms(kdnb_ctx* ctx, int begin_or_end, double* ms_out) {
  CTX_OR_FAIL(ctx);
  if (begin_or_end == 0) {
    KDNB_CUDA_TRY(c, cudaEventRecord(c->sw_begin, c->stream));
    return 0;
  }
  KDNB_CUDA_TRY(c, cudaEventRecord(c->sw_end, c->stream));
  KDNB_CUDA_TRY(c, cudaEventSynchronize(c->sw_end));
  float t = 0.f;
  KDNB_CUDA_TRY(c, cudaEventElapsedTime(&t, c->sw_begin, c->sw_end));
  if (ms_out) *ms_out = t;
  return 0;
}

void* kdnb_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}

void kdnb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
