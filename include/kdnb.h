/*
 * kdnb.h — C ABI of libkdnb.so: the kD-tree N-body step of MarkCLewis/MultiLanguageKDTree
 * (Parallel/RustVersion) on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE path of the reference: `simple_sim` and the three stages it
 * calls (tree build, theta-criterion walk, kick/drift).  The reference has no FFI of its own
 * (SURVEY.md §8b); every entry point below cites the Rust item it replaces, relative to
 * /root/reference/Parallel/RustVersion/src/.  INTEGRATION.md shows the Rust `extern "C"` block and the
 * patched `simple_sim` a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success or a negative KDNB_E_* code
 * and never aborts the host process (the reference panics instead); the message is available from
 * kdnb_last_error().  There is NO CPU fallback: without a CUDA device kdnb_create() fails.
 * One context drives one GPU from one host thread; multi-GPU = one process/context per GPU joined with
 * kdnb_comm_init().
 */
#ifndef KDNB_H
#define KDNB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KDNB_VERSION 100

enum {
  KDNB_OK = 0,
  KDNB_E_INVALID = -1,  /* bad argument / call order */
  KDNB_E_CUDA = -2,     /* CUDA runtime error (message in kdnb_last_error) */
  KDNB_E_NOMEM = -3,    /* device or host allocation failed */
  KDNB_E_NCCL = -4,     /* NCCL error or libnccl not loadable */
  KDNB_E_CAPACITY = -5  /* caller buffer too small */
};

/* mirrors `pub struct Particle` (array_particle.rs:3-8): same field order, 64 bytes */
typedef struct kdnb_particle {
  double p[3];
  double v[3];
  double r;
  double m;
} kdnb_particle;

/* mirrors the Sequential crate's SIMD particle, `pub struct Particle { p: f64x4, v: f64x4, r: f64, m: f64 }`
 * (Sequential/RustVersion/src/simd_particle.rs:3-8): f64x4 is 32-byte aligned, so the record is 96 bytes.  Lane 3 of p
 * and v is padding: the reference's own generators leave it 0 (simd_particle.rs:13-26, :30-52) and its arithmetic keeps
 * it 0 (`(d * d).reduce_sum()`, simd_kd_tree.rs:151-152, adds an exact 0).  A non-zero lane 3 would be a fourth spatial
 * dimension there; the kdnb_*_simd calls reject it (KDNB_E_INVALID at the next synchronising call). */
typedef struct kdnb_particle_simd {
  double p[4];
  double v[4];
  double r;
  double m;
  double pad_[2];
} kdnb_particle_simd;

enum { KDNB_LEAF = 0, KDNB_INTERNAL = 1 };
#define KDNB_NO_INDEX UINT64_MAX /* usize::MAX of NEGS (array_kd_tree.rs:16) */

/* mirrors `pub enum KDTree` (array_kd_tree.rs:18-34) as a flat record.
 *   Leaf     { num_parts, leaf_parts }  -> kind=KDNB_LEAF, num_parts, leaf_parts = indices[leaf_first .. leaf_first+num_parts)
 *   Internal { split_dim, split_val, m, cm, size, left, right } -> kind=KDNB_INTERNAL and those fields
 * A slot the build never wrote (the padded layout leaves about half of them) is the reference's default
 * Leaf{0, NEGS}: kind=KDNB_LEAF, num_parts=0, leaf_first=KDNB_NO_INDEX. */
typedef struct kdnb_node {
  uint32_t kind;
  uint32_t split_dim;
  uint64_t num_parts;
  uint64_t leaf_first;
  double split_val;
  double m;
  double cm[3];
  double size;
  uint64_t left;
  uint64_t right;
} kdnb_node;

enum {
  KDNB_LAYOUT_PADDED = 0, /* build_tree_par4 (array_kd_tree.rs:515-583): right = cur+1+nodes_needed(left_len) */
  KDNB_LAYOUT_DENSE = 1   /* build_tree (array_kd_tree.rs:63-130): right = last node of left subtree + 1 */
};

enum {
  KDNB_FLAG_PROFILE = 1u,     /* record CUDA events around build / walk / kick / exchange */
  KDNB_FLAG_WALK_COUNTS = 2u, /* walk also counts node tests / accepts / leaf visits / pair interactions per particle */
  KDNB_FLAG_EXACT_MATH = 4u   /* force magnitudes with IEEE sqrt and divide exactly as the reference writes them
                                 (default: rsqrt-based, <= 2 ulp apart; acceptance tests are always exact).
                                 Where a squared pair distance is 0 or underflows to 0 (coincident particles,
                                 separations below ~1e-154) the reference divides by zero and its acceleration is
                                 non-finite; the default path is non-finite for exactly the same particles but may
                                 hold NaN where the reference holds +-inf — with this flag the values are identical */
};

typedef struct kdnb_config {
  uint32_t struct_size; /* = sizeof(kdnb_config) */
  int32_t device;       /* CUDA device ordinal */
  uint32_t max_parts;   /* MAX_PARTS (array_kd_tree.rs:14), 4..32; 0 -> 8 */
  int32_t layout;       /* KDNB_LAYOUT_* */
  double theta;         /* THETA (array_kd_tree.rs:15); 0 -> 0.3 */
  uint32_t flags;       /* KDNB_FLAG_* */
  uint32_t reserved;
} kdnb_config;

typedef struct kdnb_ctx kdnb_ctx;

/* ---- lifetime: the prologue of simple_sim (array_kd_tree.rs:624-630) owns acc / tree / indices; here the context does */
kdnb_ctx* kdnb_create(const kdnb_config* cfg); /* NULL on failure: see kdnb_last_error(NULL) */
void kdnb_destroy(kdnb_ctx* ctx);
const char* kdnb_last_error(const kdnb_ctx* ctx); /* ctx may be NULL (creation errors) */
int kdnb_version(void);

/* ---- state: `bodies: &mut Vec<Particle>` (array_kd_tree.rs:623).  AoS in, AoS out, original order kept.
 * Uploads enqueue their host-to-device copy and return: a page-locked source (kdnb_host_alloc) must stay untouched until
 * kdnb_synchronize or a download returns (pageable memory is staged by the driver before the call returns). */
int kdnb_upload_particles(kdnb_ctx* ctx, const kdnb_particle* aos, uint64_t count);
int kdnb_download_particles(kdnb_ctx* ctx, kdnb_particle* out, uint64_t capacity);
uint64_t kdnb_particle_count(const kdnb_ctx* ctx);

/* ---- the three stages of one step */
int kdnb_build_tree(kdnb_ctx* ctx);             /* indices reset + build_tree_par4 / build_tree (array_kd_tree.rs:641-643) */
int kdnb_calc_accel(kdnb_ctx* ctx);             /* acc[i] = calc_accel(i, bodies, tree) for all i (array_kd_tree.rs:647, :585-621) */
int kdnb_kick_drift(kdnb_ctx* ctx, double dt);  /* v += dt*a; p += dt*v; a = 0 (array_kd_tree.rs:649-662) */

/* ---- the driver: simple_sim(bodies, dt, steps) (array_kd_tree.rs:623-664) */
int kdnb_simple_sim(kdnb_ctx* ctx, double dt, int64_t steps); /* on the uploaded state; asynchronous until a download / kdnb_synchronize */
/* one-call drop-in with host buffers: upload, `steps` steps, download into `bodies` */
int kdnb_simple_sim_host(const kdnb_config* cfg, kdnb_particle* bodies, uint64_t count, double dt, int64_t steps);
/* the same on an existing context (buffers are reused across calls): upload + `steps` steps + download */
int kdnb_simple_sim_bodies(kdnb_ctx* ctx, kdnb_particle* bodies, uint64_t count, double dt, int64_t steps);
int kdnb_synchronize(kdnb_ctx* ctx); /* also reports what only the device saw: a build look-back or (multi-GPU) a peer that
                                       never published its accelerations within ~15 s; so do the particle downloads */

/* ---- results of the stages (parity hooks) */
int kdnb_download_accel(kdnb_ctx* ctx, double* acc /* count*3, original particle order */);
int kdnb_upload_accel(kdnb_ctx* ctx, const double* acc /* count*3, original order; test hook for kick_drift */);
/* `tree: Vec<KDTree>` + `indices` after the build.  nodes: capacity `cap` records, receives kdnb_node_count();
 * indices: count entries (tree order), may be NULL.  Node index positions are exactly the reference layout's. */
int kdnb_download_tree(kdnb_ctx* ctx, kdnb_node* nodes, uint64_t cap, uint64_t* n_nodes, uint64_t* indices);
/* per particle {internal nodes tested, monopoles accepted, leaves visited, pair interactions}; needs KDNB_FLAG_WALK_COUNTS */
int kdnb_download_walk_counts(kdnb_ctx* ctx, uint64_t* counts /* count*4, original order */);

/* ---- quickstat_index (Parallel/RustVersion/src/quickstat.rs:9-34; stand-alone bench src/bin/bench_quickstat.rs).
 * Permutes indices[0..count) (element ids into vals[0..n_vals)) so that indices[goal] refers to the goal-th smallest
 * value, nothing before it is larger and nothing after it is smaller (the post-condition of quickstat.rs:199-253).
 * Device radix select + one stable three-way partition: of the permutations the reference's random pivots can
 * produce, the one that keeps the input order inside {< pivot}, {== pivot}, {> pivot}.  Host arrays in, host arrays
 * out; *device_ms (may be NULL) receives the device time of the selection without the copies.
 * Where the reference would panic (goal >= count, an index >= n_vals) a negative code is returned. */
int kdnb_quickstat_index(kdnb_ctx* ctx, const double* vals, uint64_t n_vals, uint64_t* indices, uint64_t count,
                         uint64_t goal, double* device_ms);

/* ---- pure functions */
uint64_t kdnb_nodes_needed(uint64_t num_parts, uint32_t max_parts); /* nodes_needed_for_particles (array_kd_tree.rs:45-53) */
uint64_t kdnb_node_count(const kdnb_ctx* ctx); /* allocate_node_vec(count).len() for this context's layout (array_kd_tree.rs:55-60) */

/* ---- multi-GPU (new; the reference is single-process).  Tree replicated, walk sharded by tree-ordered
 * ranges, tree-ordered accelerations exchanged with one ncclAllGather per step. */
/* pure: the tree-slot range [begin, end) rank `rank` of `world_size` walks for `count` particles (warp-aligned shards) */
int kdnb_shard_range(uint64_t count, int rank, int world_size, uint64_t* begin, uint64_t* end);
/* multi-GPU, sharded tree build: what rank `rank` of a power-of-two `world_size` builds below the top log2(world_size)
 * levels — the subtree of its level-log2(world) segment: tree slots [first_slot, first_slot + slots) and node indices
 * [first_node, first_node + nodes) in the given layout (closed forms of `count`: array_kd_tree.rs:560, :566-581,
 * :120-126).  Pure; KDNB_E_INVALID when world_size is not a power of two. */
int kdnb_build_shard_plan(uint64_t count, uint32_t max_parts, int layout, int rank, int world_size, uint64_t* first_slot,
                          uint64_t* slots, uint64_t* first_node, uint64_t* nodes);
int kdnb_comm_unique_id(void* id_out_128_bytes);
int kdnb_comm_init(kdnb_ctx* ctx, const void* id_128_bytes, int rank, int world_size);

/* Sharded host buffers for multi-GPU callers: rank r owns particles [first, first+count) of the global array
 * (kdnb_host_shard_range; equal shards of ceil(total/world)).  Upload sends only the own shard over PCIe and
 * all-gathers the rest over NVLink; download returns only the own shard (every replica holds the full state). */
int kdnb_host_shard_range(uint64_t total, int rank, int world_size, uint64_t* first, uint64_t* count);
int kdnb_upload_particles_sharded(kdnb_ctx* ctx, const kdnb_particle* shard, uint64_t total);
int kdnb_download_particles_sharded(kdnb_ctx* ctx, kdnb_particle* shard_out);
int kdnb_simple_sim_bodies_sharded(kdnb_ctx* ctx, kdnb_particle* shard, uint64_t total, double dt, int64_t steps);

/* ---- measurement */
enum { KDNB_STAGE_BUILD = 0, KDNB_STAGE_WALK = 1, KDNB_STAGE_KICK = 2, KDNB_STAGE_EXCHANGE = 3, KDNB_STAGE_COUNT = 4 };
int kdnb_stage_ms(kdnb_ctx* ctx, double ms_out[KDNB_STAGE_COUNT], uint64_t* steps_out); /* sums since last reset; needs KDNB_FLAG_PROFILE; synchronizes */
int kdnb_stage_reset(kdnb_ctx* ctx);
uint64_t kdnb_launch_count(const kdnb_ctx* ctx);     /* kernels launched by this context so far */
int kdnb_measure_fp64_peak(kdnb_ctx* ctx, double* tflops_out); /* DFMA-chain microbenchmark: the walk's roofline denominator */
int kdnb_flush_l2(kdnb_ctx* ctx);                     /* overwrite a 256 MiB scratch buffer */
int kdnb_device_ms(kdnb_ctx* ctx, int begin_or_end, double* ms_out); /* CUDA-event stopwatch on the context's stream */

/* ---- the Sequential crate's SIMD particle surface (simd_particle.rs / simd_kd_tree.rs:169-202): the same step on
 * 96-byte f64x4 records; create the context with max_parts = 7 and KDNB_LAYOUT_DENSE for that crate's tree
 * (simd_kd_tree.rs:9, :49-138).  Same arithmetic as the scalar record: the results are bit-identical. */
int kdnb_upload_particles_simd(kdnb_ctx* ctx, const kdnb_particle_simd* aos, uint64_t count);
int kdnb_download_particles_simd(kdnb_ctx* ctx, kdnb_particle_simd* out, uint64_t capacity);
int kdnb_simple_sim_bodies_simd(kdnb_ctx* ctx, kdnb_particle_simd* bodies, uint64_t count, double dt, int64_t steps);

/* page-locked host buffers for the transfers above (optional; any host pointer works, pinned is faster) */
void* kdnb_host_alloc(uint64_t bytes);
void kdnb_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* KDNB_H */
