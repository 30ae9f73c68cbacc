/*
 * kdtree_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See kdtree_oracle.h.
 *
 * Restates, operation for operation, the reference's Rust hot path.  Compile with
 *   gcc -O2 -ffp-contract=off -fopenmp   (never -Ofast / -ffast-math: rustc never contracts or
 *   reassociates f64 arithmetic, and parity below is bit-level).
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/Parallel/RustVersion/src/).
 */
#include "kdtree_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ RNG (replaces fastrand) */

uint64_t okd_rng_next(uint64_t* state) {
  uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

int okd_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

uint64_t okd_sizeof_node(void) { return (uint64_t)sizeof(okd_node); }

/* ------------------------------------------------------------------ particles */

/* array_particle.rs:10-17 */
void okd_two_bodies(okd_particle out[2]) {
  memset(out, 0, 2 * sizeof(okd_particle));
  out[0].r = 1.0;
  out[0].m = 1.0;
  out[1].p[0] = 1.0;
  out[1].v[1] = 1.0;
  out[1].r = 1e-4;
  out[1].m = 1e-20;
}

/* array_particle.rs:19-44.  fastrand::f64() (:31) is unseeded in the reference; here the
 * uniform deviate is the top 53 bits of splitmix64(seed). */
void okd_circular_orbits(uint64_t n, uint64_t seed, okd_particle* out) {
  uint64_t st = seed;
  memset(&out[0], 0, sizeof(okd_particle));
  out[0].r = 0.00465047;
  out[0].m = 1.0;
  for (uint64_t i = 0; i < n; ++i) {
    double d = 0.1 + ((double)i * 5.0 / (double)n);
    double v = sqrt(1.0 / d);
    double u = (double)(okd_rng_next(&st) >> 11) * 0x1.0p-53;
    double theta = u * 6.28;
    double x = d * cos(theta);
    double y = d * sin(theta);
    double vx = -v * sin(theta);
    double vy = v * cos(theta);
    okd_particle* q = &out[i + 1];
    q->p[0] = x;
    q->p[1] = y;
    q->p[2] = 0.0;
    q->v[0] = vx;
    q->v[1] = vy;
    q->v[2] = 0.0;
    q->m = 1e-14;
    q->r = 1e-7;
  }
}

/* array_particle.rs:67-76 */
static inline void calc_pp_accel(const okd_particle* pi, const okd_particle* pj, double out[3]) {
  double dx = pi->p[0] - pj->p[0];
  double dy = pi->p[1] - pj->p[1];
  double dz = pi->p[2] - pj->p[2];
  double dp2 = dx * dx + dy * dy + dz * dz;
  double dist = sqrt(dp2);
  double magi = -pj->m / (dist * dist * dist);
  out[0] = magi * dx;
  out[1] = magi * dy;
  out[2] = magi * dz;
}

/* ------------------------------------------------------------------ allocation */

static uint64_t ceil_log2_u64(uint64_t k) { /* ceil(log2(k)) for k >= 1, as f64::log2(k).ceil() */
  uint64_t e = 0;
  while ((1ull << e) < k) ++e;
  return e;
}

/* array_kd_tree.rs:45-53 */
uint64_t okd_nodes_needed_for_particles(uint64_t num_parts, uint64_t max_parts) {
  if (num_parts <= max_parts) return 1;
  uint64_t min_num_leaves = num_parts / (max_parts / 2);
  uint64_t num_leaves = 1ull << ceil_log2_u64(min_num_leaves);
  return 2 * num_leaves - 1;
}

/* array_kd_tree.rs:16,55-60 — Leaf{0, [usize::MAX; MAX_PARTS]} */
void okd_fill_default_nodes(okd_node* nodes, uint64_t count) {
  for (uint64_t i = 0; i < count; ++i) {
    memset(&nodes[i], 0, sizeof(okd_node));
    nodes[i].is_internal = 0;
    nodes[i].u.leaf.num_parts = 0;
    for (int k = 0; k < OKD_LEAF_CAP; ++k) nodes[i].u.leaf.leaf_parts[k] = UINT64_MAX;
  }
}

/* ------------------------------------------------------------------ quickstat */

/* quickstat.rs:9-34, comparator |a,b| key(a) < key(b) with key = stride-addressed f64 */
static void quickstat_strided(uint64_t* indices, uint64_t len, uint64_t goal, const char* base,
                              size_t stride, uint64_t* rng) {
#define KEY(i) (*(const double*)(base + (size_t)(i) * stride))
  uint64_t s = 0, e = len;
  while (s + 1 < e) {
    uint64_t pivot = s + okd_rng_next(rng) % (e - s);
    uint64_t t = indices[s];
    indices[s] = indices[pivot];
    indices[pivot] = t;
    uint64_t low = s + 1, high = e - 1;
    while (low <= high) {
      if (KEY(indices[low]) < KEY(indices[s])) {
        low += 1;
      } else {
        t = indices[low];
        indices[low] = indices[high];
        indices[high] = t;
        high -= 1;
      }
    }
    t = indices[s];
    indices[s] = indices[high];
    indices[high] = t;
    if (high < goal) {
      s = high + 1;
    } else if (high > goal) {
      e = high;
    } else {
      s = e;
    }
  }
#undef KEY
}

void okd_quickstat_index_f64(uint64_t* indices, uint64_t len, uint64_t goal, const double* vals,
                             uint64_t* rng_state) {
  quickstat_strided(indices, len, goal, (const char*)vals, sizeof(double), rng_state);
}

/* ------------------------------------------------------------------ node statistics */

typedef struct node_stats {
  double m, cm[3], min[3], max[3], size;
  uint64_t split_dim;
} node_stats;

/* array_kd_tree.rs:532-557 (== :83-108) — sequential scan in current slice order */
static void scan_stats(const uint64_t* indices, uint64_t len, const okd_particle* P, node_stats* s) {
  double mn[3] = {1e100, 1e100, 1e100};
  double mx[3] = {-1e100, -1e100, -1e100};
  double m = 0.0;
  double cm[3] = {0.0, 0.0, 0.0};
  for (uint64_t i = 0; i < len; ++i) {
    const okd_particle* q = &P[indices[i]];
    m += q->m;
    cm[0] += q->m * q->p[0];
    cm[1] += q->m * q->p[1];
    cm[2] += q->m * q->p[2];
    mn[0] = fmin(mn[0], q->p[0]);
    mn[1] = fmin(mn[1], q->p[1]);
    mn[2] = fmin(mn[2], q->p[2]);
    mx[0] = fmax(mx[0], q->p[0]);
    mx[1] = fmax(mx[1], q->p[1]);
    mx[2] = fmax(mx[2], q->p[2]);
  }
  cm[0] /= m;
  cm[1] /= m;
  cm[2] /= m;
  uint64_t split_dim = 0;
  for (uint64_t dim = 1; dim < 3; ++dim) {
    if (mx[dim] - mn[dim] > mx[split_dim] - mn[split_dim]) split_dim = dim;
  }
  s->m = m;
  for (int k = 0; k < 3; ++k) {
    s->cm[k] = cm[k];
    s->min[k] = mn[k];
    s->max[k] = mx[k];
  }
  s->split_dim = split_dim;
  s->size = mx[split_dim] - mn[split_dim];
}

static void write_leaf(okd_node* node, const uint64_t* indices, uint64_t np) {
  memset(node, 0, sizeof(okd_node));
  node->is_internal = 0;
  node->u.leaf.num_parts = np;
  for (uint64_t i = 0; i < OKD_LEAF_CAP; ++i) node->u.leaf.leaf_parts[i] = 0; /* `[0; MAX_PARTS]` :72,:525 */
  for (uint64_t i = 0; i < np; ++i) node->u.leaf.leaf_parts[i] = indices[i];
}

static void write_internal(okd_node* node, const node_stats* s, double split_val, uint64_t left,
                           uint64_t right) {
  node->is_internal = 1;
  node->u.in.split_dim = s->split_dim;
  node->u.in.split_val = split_val;
  node->u.in.m = s->m;
  node->u.in.cm[0] = s->cm[0];
  node->u.in.cm[1] = s->cm[1];
  node->u.in.cm[2] = s->cm[2];
  node->u.in.size = s->size;
  node->u.in.left = left;
  node->u.in.right = right;
}

/* ------------------------------------------------------------------ build_tree (dense) */

/* array_kd_tree.rs:63-130 */
uint64_t okd_build_tree(uint64_t* indices, uint64_t start, uint64_t end, const okd_particle* particles,
                        uint64_t cur_node, okd_node* nodes, uint64_t cap, uint64_t max_parts,
                        uint64_t* rng_state) {
  uint64_t np = end - start;
  if (cur_node >= cap) return UINT64_MAX; /* the Rust code would `resize` (:75,:123) */
  if (np <= max_parts) {
    write_leaf(&nodes[cur_node], indices + start, np);
    return cur_node;
  }
  node_stats st;
  scan_stats(indices + start, np, particles, &st);
  uint64_t mid = (start + end) / 2;
  quickstat_strided(indices + start, np, mid - start, (const char*)&particles[0].p[st.split_dim],
                    sizeof(okd_particle), rng_state);
  double split_val = particles[indices[mid]].p[st.split_dim];
  uint64_t left = okd_build_tree(indices, start, mid, particles, cur_node + 1, nodes, cap, max_parts, rng_state);
  if (left == UINT64_MAX) return UINT64_MAX;
  uint64_t right = okd_build_tree(indices, mid, end, particles, left + 1, nodes, cap, max_parts, rng_state);
  if (right == UINT64_MAX) return UINT64_MAX;
  write_internal(&nodes[cur_node], &st, split_val, cur_node + 1, left + 1);
  return right;
}

/* ------------------------------------------------------------------ build_tree_par4 (padded) */

static uint64_t node_seed(uint64_t seed, uint64_t cur_node) {
  uint64_t s = seed ^ (cur_node * 0xD6E8FEB86659FD93ull);
  (void)okd_rng_next(&s);
  return s;
}

/* array_kd_tree.rs:515-583 */
static void par4_rec(uint64_t* indices, uint64_t np, uint64_t cur_node, const okd_particle* particles,
                     okd_node* nodes, uint64_t max_parts, uint64_t seed, uint64_t thread_cnt,
                     uint64_t max_threads) {
  if (np <= max_parts) {
    write_leaf(&nodes[0], indices, np);
    return;
  }
  node_stats st;
  scan_stats(indices, np, particles, &st);
  uint64_t mid = np / 2;
  uint64_t rng = node_seed(seed, cur_node);
  quickstat_strided(indices, np, mid, (const char*)&particles[0].p[st.split_dim], sizeof(okd_particle), &rng);
  double split_val = particles[indices[mid]].p[st.split_dim];
  uint64_t num_nodes = okd_nodes_needed_for_particles(mid, max_parts);
  okd_node* left_nodes = nodes + 1;
  okd_node* right_nodes = nodes + 1 + num_nodes;
  if (thread_cnt < max_threads) { /* rayon::join :572-575 */
#pragma omp task default(shared)
    par4_rec(indices, mid, cur_node + 1, particles, left_nodes, max_parts, seed, thread_cnt * 2, max_threads);
#pragma omp task default(shared)
    par4_rec(indices + mid, np - mid, cur_node + 1 + num_nodes, particles, right_nodes, max_parts, seed,
             thread_cnt * 2, max_threads);
#pragma omp taskwait
  } else {
    par4_rec(indices, mid, cur_node + 1, particles, left_nodes, max_parts, seed, thread_cnt * 2, max_threads);
    par4_rec(indices + mid, np - mid, cur_node + 1 + num_nodes, particles, right_nodes, max_parts, seed,
             thread_cnt * 2, max_threads);
  }
  write_internal(&nodes[0], &st, split_val, cur_node + 1, cur_node + 1 + num_nodes);
}

void okd_build_tree_par4(uint64_t* indices, uint64_t len, uint64_t cur_node, const okd_particle* particles,
                         okd_node* nodes, uint64_t max_parts, uint64_t seed, int max_threads) {
  if (max_threads <= 1) {
    par4_rec(indices, len, cur_node, particles, nodes, max_parts, seed, 1, 1);
    return;
  }
#pragma omp parallel num_threads(max_threads)
#pragma omp single
  par4_rec(indices, len, cur_node, particles, nodes, max_parts, seed, 1, (uint64_t)max_threads);
}

/* ------------------------------------------------------------------ canonical build */

/* number of nodes of the dense subtree over `len` particles (what build_tree :63-130 consumes) */
static uint64_t dense_subtree_nodes(uint64_t len, uint64_t mp) {
  if (len <= mp) return 1;
  uint64_t half = len / 2;
  return 1 + dense_subtree_nodes(half, mp) + dense_subtree_nodes(len - half, mp);
}

typedef struct canon_ctx {
  const okd_particle* P;
  okd_node* nodes;
  uint64_t cap, mp;
  int layout;
  uint64_t max_threads;
  int overflow;
} canon_ctx;

static inline int key_less(const okd_particle* P, uint64_t a, uint64_t b, uint64_t sd) {
  double ka = P[a].p[sd], kb = P[b].p[sd];
  return (ka < kb) || (ka == kb && a < b); /* canonical total order: ties -> lower particle index */
}

/* deterministic quick-select under the total order (result sets are unique, pivots irrelevant) */
static void canon_select(uint64_t* idx, uint64_t len, uint64_t goal, const okd_particle* P, uint64_t sd) {
  uint64_t s = 0, e = len;
  uint64_t rng = 0x1234567ull + len;
  while (s + 1 < e) {
    uint64_t pivot = s + okd_rng_next(&rng) % (e - s);
    uint64_t t = idx[s];
    idx[s] = idx[pivot];
    idx[pivot] = t;
    uint64_t low = s + 1, high = e - 1;
    while (low <= high) {
      if (key_less(P, idx[low], idx[s], sd)) {
        low += 1;
      } else {
        t = idx[low];
        idx[low] = idx[high];
        idx[high] = t;
        high -= 1;
      }
    }
    t = idx[s];
    idx[s] = idx[high];
    idx[high] = t;
    if (high < goal) s = high + 1;
    else if (high > goal) e = high;
    else s = e;
  }
}

static int cmp_u64(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return (x > y) - (x < y);
}

/* ms = {M, Sx, Sy, Sz}: leaf = sequential over ascending ids; internal = left + right */
static void canon_rec(canon_ctx* c, uint64_t* idx, uint64_t len, uint64_t cur, double ms[4],
                      uint64_t thread_cnt) {
  if (cur >= c->cap) {
    c->overflow = 1;
    ms[0] = ms[1] = ms[2] = ms[3] = 0.0;
    return;
  }
  const okd_particle* P = c->P;
  if (len <= c->mp) {
    qsort(idx, len, sizeof(uint64_t), cmp_u64);
    write_leaf(&c->nodes[cur], idx, len);
    double m = 0.0, s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (uint64_t i = 0; i < len; ++i) {
      const okd_particle* q = &P[idx[i]];
      m += q->m;
      s0 += q->m * q->p[0];
      s1 += q->m * q->p[1];
      s2 += q->m * q->p[2];
    }
    ms[0] = m;
    ms[1] = s0;
    ms[2] = s1;
    ms[3] = s2;
    return;
  }
  double mn[3] = {1e100, 1e100, 1e100}, mx[3] = {-1e100, -1e100, -1e100};
  for (uint64_t i = 0; i < len; ++i) {
    const okd_particle* q = &P[idx[i]];
    for (int k = 0; k < 3; ++k) {
      mn[k] = fmin(mn[k], q->p[k]);
      mx[k] = fmax(mx[k], q->p[k]);
    }
  }
  node_stats st;
  memset(&st, 0, sizeof st);
  st.split_dim = 0;
  for (uint64_t dim = 1; dim < 3; ++dim)
    if (mx[dim] - mn[dim] > mx[st.split_dim] - mn[st.split_dim]) st.split_dim = dim;
  st.size = mx[st.split_dim] - mn[st.split_dim];
  uint64_t mid = len / 2;
  canon_select(idx, len, mid, P, st.split_dim);
  double split_val = P[idx[mid]].p[st.split_dim];
  uint64_t left = cur + 1;
  uint64_t right = (c->layout == OKD_LAYOUT_PADDED) ? cur + 1 + okd_nodes_needed_for_particles(mid, c->mp)
                                                    : cur + 1 + dense_subtree_nodes(mid, c->mp);
  double lm[4], rm[4];
  if (thread_cnt < c->max_threads) {
#pragma omp task default(shared)
    canon_rec(c, idx, mid, left, lm, thread_cnt * 2);
#pragma omp task default(shared)
    canon_rec(c, idx + mid, len - mid, right, rm, thread_cnt * 2);
#pragma omp taskwait
  } else {
    canon_rec(c, idx, mid, left, lm, thread_cnt * 2);
    canon_rec(c, idx + mid, len - mid, right, rm, thread_cnt * 2);
  }
  for (int k = 0; k < 4; ++k) ms[k] = lm[k] + rm[k];
  st.m = ms[0];
  st.cm[0] = ms[1] / ms[0];
  st.cm[1] = ms[2] / ms[0];
  st.cm[2] = ms[3] / ms[0];
  write_internal(&c->nodes[cur], &st, split_val, left, right);
}

static uint64_t last_used_node(const okd_node* nodes, uint64_t root) {
  uint64_t cur = root;
  while (nodes[cur].is_internal) cur = nodes[cur].u.in.right;
  return cur;
}

uint64_t okd_build_tree_canonical(uint64_t* indices, uint64_t n, const okd_particle* particles,
                                  okd_node* nodes, uint64_t cap, uint64_t max_parts, int layout,
                                  int max_threads) {
  canon_ctx c = {particles, nodes, cap, max_parts, layout, (uint64_t)(max_threads < 1 ? 1 : max_threads), 0};
  double ms[4];
  if (max_threads <= 1) {
    canon_rec(&c, indices, n, 0, ms, 1);
  } else {
#pragma omp parallel num_threads(max_threads)
#pragma omp single
    canon_rec(&c, indices, n, 0, ms, 1);
  }
  if (c.overflow) return UINT64_MAX;
  return last_used_node(nodes, 0);
}

/* ------------------------------------------------------------------ walk */

/* array_kd_tree.rs:585-617 */
static void accel_recur(uint64_t cur_node, uint64_t p, const okd_particle* particles, const okd_node* nodes,
                        double theta2, double out[3], okd_walk_counts* cnt) {
  const okd_node* nd = &nodes[cur_node];
  if (!nd->is_internal) {
    double acc[3] = {0.0, 0.0, 0.0};
    if (cnt) cnt->leaf_visits += 1;
    for (uint64_t i = 0; i < nd->u.leaf.num_parts; ++i) {
      if (nd->u.leaf.leaf_parts[i] != p) {
        double pp[3];
        calc_pp_accel(&particles[p], &particles[nd->u.leaf.leaf_parts[i]], pp);
        acc[0] += pp[0];
        acc[1] += pp[1];
        acc[2] += pp[2];
        if (cnt) cnt->pp += 1;
      }
    }
    out[0] = acc[0];
    out[1] = acc[1];
    out[2] = acc[2];
    return;
  }
  double dx = particles[p].p[0] - nd->u.in.cm[0];
  double dy = particles[p].p[1] - nd->u.in.cm[1];
  double dz = particles[p].p[2] - nd->u.in.cm[2];
  double dist_sqr = dx * dx + dy * dy + dz * dz;
  if (cnt) cnt->node_visits += 1;
  if (nd->u.in.size * nd->u.in.size < theta2 * dist_sqr) {
    double dist = sqrt(dist_sqr);
    double magi = -nd->u.in.m / (dist_sqr * dist);
    out[0] = dx * magi;
    out[1] = dy * magi;
    out[2] = dz * magi;
    if (cnt) cnt->accepts += 1;
  } else {
    double l[3], r[3];
    accel_recur(nd->u.in.left, p, particles, nodes, theta2, l, cnt);
    accel_recur(nd->u.in.right, p, particles, nodes, theta2, r, cnt);
    out[0] = l[0] + r[0];
    out[1] = l[1] + r[1];
    out[2] = l[2] + r[2];
  }
}

/* array_kd_tree.rs:619-621; `THETA * THETA * dist_sqr` parses as (THETA*THETA)*dist_sqr (:606) */
void okd_calc_accel(uint64_t p, const okd_particle* particles, const okd_node* nodes, double theta,
                    double out[3]) {
  accel_recur(0, p, particles, nodes, theta * theta, out, NULL);
}

void okd_calc_accel_counted(uint64_t p, const okd_particle* particles, const okd_node* nodes, double theta,
                            double out[3], okd_walk_counts* counts) {
  memset(counts, 0, sizeof *counts);
  accel_recur(0, p, particles, nodes, theta * theta, out, counts);
}

void okd_calc_accel_all(uint64_t n, const okd_particle* particles, const okd_node* nodes, double theta,
                        double* acc, okd_walk_counts* counts, int max_threads) {
  double theta2 = theta * theta;
  (void)max_threads;
#pragma omp parallel for schedule(dynamic, 256) num_threads(max_threads < 1 ? 1 : max_threads)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    if (counts) {
      memset(&counts[i], 0, sizeof(okd_walk_counts));
      accel_recur(0, (uint64_t)i, particles, nodes, theta2, &acc[3 * i], &counts[i]);
    } else {
      accel_recur(0, (uint64_t)i, particles, nodes, theta2, &acc[3 * i], NULL);
    }
  }
}

/* ------------------------------------------------------------------ kick/drift + driver */

/* array_kd_tree.rs:649-662 */
void okd_kick_drift(uint64_t n, okd_particle* bodies, double* acc, double dt, int max_threads) {
  (void)max_threads;
#pragma omp parallel for schedule(static) num_threads(max_threads < 1 ? 1 : max_threads)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    okd_particle* b = &bodies[i];
    double* a = &acc[3 * i];
    b->v[0] += dt * a[0];
    b->v[1] += dt * a[1];
    b->v[2] += dt * a[2];
    double dx = dt * b->v[0];
    double dy = dt * b->v[1];
    double dz = dt * b->v[2];
    b->p[0] += dx;
    b->p[1] += dy;
    b->p[2] += dz;
    a[0] = 0.0;
    a[1] = 0.0;
    a[2] = 0.0;
  }
}

/* array_kd_tree.rs:623-664 */
int okd_simple_sim(okd_particle* bodies, uint64_t n, double dt, int64_t steps, uint64_t max_parts,
                   double theta, int layout, int order, uint64_t seed, int max_threads) {
  if (max_parts < 2 || max_parts > OKD_LEAF_CAP) return -1;
  double* acc = (double*)calloc(3 * n, sizeof(double));
  uint64_t cap = (layout == OKD_LAYOUT_PADDED) ? okd_nodes_needed_for_particles(n, max_parts)
                                               : dense_subtree_nodes(n, max_parts);
  okd_node* tree = (okd_node*)malloc(cap * sizeof(okd_node));
  uint64_t* indices = (uint64_t*)malloc(n * sizeof(uint64_t));
  if (!acc || !tree || !indices) {
    free(acc);
    free(tree);
    free(indices);
    return -2;
  }
  okd_fill_default_nodes(tree, cap);
  uint64_t rng = seed;
  int rc = 0;
  for (int64_t step = 0; step < steps; ++step) {
#pragma omp parallel for schedule(static) num_threads(max_threads < 1 ? 1 : max_threads)
    for (int64_t i = 0; i < (int64_t)n; ++i) indices[i] = (uint64_t)i; /* :641 */
    if (order == OKD_ORDER_CANONICAL) {
      if (okd_build_tree_canonical(indices, n, bodies, tree, cap, max_parts, layout, max_threads) == UINT64_MAX)
        rc = -3;
    } else if (layout == OKD_LAYOUT_PADDED) {
      okd_build_tree_par4(indices, n, 0, bodies, tree, max_parts, seed + (uint64_t)step, max_threads); /* :643 */
    } else {
      if (okd_build_tree(indices, 0, n, bodies, 0, tree, cap, max_parts, &rng) == UINT64_MAX) rc = -3;
    }
    if (rc) break;
    okd_calc_accel_all(n, bodies, tree, theta, acc, NULL, max_threads); /* :647 */
    okd_kick_drift(n, bodies, acc, dt, max_threads);                    /* :649-662 */
  }
  free(acc);
  free(tree);
  free(indices);
  return rc;
}

/* ------------------------------------------------------------------ invariant + dump */

static uint64_t check_rec(uint64_t node, const okd_node* nodes, const okd_particle* P, double mn[3],
                          double mx[3], int dims) {
  const okd_node* nd = &nodes[node];
  if (!nd->is_internal) {
    for (uint64_t k = 0; k < nd->u.leaf.num_parts; ++k) {
      uint64_t i = nd->u.leaf.leaf_parts[k];
      for (int d = 0; d < dims; ++d) {
        if (!(P[i].p[d] >= mn[d])) return 1 + node;
        if (!(P[i].p[d] < mx[d])) return 1 + node;
      }
    }
    return 0;
  }
  uint64_t sd = nd->u.in.split_dim;
  double tmin = mn[sd], tmax = mx[sd];
  mx[sd] = nd->u.in.split_val;
  uint64_t r = check_rec(nd->u.in.left, nodes, P, mn, mx, dims);
  if (r) return r;
  mx[sd] = tmax;
  mn[sd] = nd->u.in.split_val;
  r = check_rec(nd->u.in.right, nodes, P, mn, mx, dims);
  mn[sd] = tmin;
  return r;
}

/* array_kd_tree.rs:834-877 (reference checks dims 0..2 only, :845) */
uint64_t okd_check_tree_struct(const okd_node* nodes, const okd_particle* particles, int dims_checked) {
  double mn[3] = {-1e100, -1e100, -1e100}, mx[3] = {1e100, 1e100, 1e100};
  return check_rec(0, nodes, particles, mn, mx, dims_checked);
}

/* array_kd_tree.rs:666-692.  Numbers are printed with %.17g (round-trippable); Rust's `{}` prints the
 * shortest round-trippable form, so files agree as numbers, not necessarily as text. */
int okd_print_tree(const char* path, const okd_node* nodes, uint64_t n_nodes, const okd_particle* particles) {
  FILE* f = fopen(path, "w");
  if (!f) return -1;
  fprintf(f, "%llu\n", (unsigned long long)n_nodes);
  for (uint64_t i = 0; i < n_nodes; ++i) {
    const okd_node* nd = &nodes[i];
    if (!nd->is_internal) {
      fprintf(f, "L %llu\n", (unsigned long long)nd->u.leaf.num_parts);
      for (uint64_t k = 0; k < nd->u.leaf.num_parts; ++k) {
        const okd_particle* q = &particles[nd->u.leaf.leaf_parts[k]];
        fprintf(f, "%.17g %.17g %.17g\n", q->p[0], q->p[1], q->p[2]);
      }
    } else {
      fprintf(f, "I %llu %.17g %llu %llu\n", (unsigned long long)nd->u.in.split_dim, nd->u.in.split_val,
              (unsigned long long)nd->u.in.left, (unsigned long long)nd->u.in.right);
    }
  }
  fclose(f);
  return 0;
}
