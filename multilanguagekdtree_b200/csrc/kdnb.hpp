// kdnb.hpp — header-only C++ host mirror of the reference's Rust modules over the C ABI (include/kdnb.h).
//
// The reference's host language is Rust; this image has no rustc/cargo, so the host side above the C ABI is C++
// (the Rust binding a maintainer would add is in INTEGRATION.md and rust/kdnb-sys/, source only).  Names, argument
// meaning and in-place semantics follow Parallel/RustVersion/src/{array_particle,array_kd_tree}.rs; where the Rust
// code panics this throws std::runtime_error.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kdnb.h"

namespace array_particle {

using Particle = kdnb_particle;  // array_particle.rs:3-8

inline uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// array_particle.rs:10-17
inline std::vector<Particle> two_bodies() {
  std::vector<Particle> b(2);
  b[0] = Particle{{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, 1.0, 1.0};
  b[1] = Particle{{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, 1e-4, 1e-20};
  return b;
}

// array_particle.rs:19-44 — n+1 particles; angles from a seeded splitmix64 stream (fastrand::f64() there, :31)
inline std::vector<Particle> circular_orbits(size_t n, uint64_t seed = 12345) {
  std::vector<Particle> buf;
  buf.reserve(n + 1);
  buf.push_back(Particle{{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, 0.00465047, 1.0});
  uint64_t st = seed;
  for (size_t i = 0; i < n; ++i) {
    const double d = 0.1 + ((double)i * 5.0 / (double)n);
    const double v = std::sqrt(1.0 / d);
    const double theta = (double)(splitmix64(st) >> 11) * 0x1.0p-53 * 6.28;
    const double x = d * std::cos(theta), y = d * std::sin(theta);
    const double vx = -v * std::sin(theta), vy = v * std::cos(theta);
    buf.push_back(Particle{{x, y, 0.0}, {vx, vy, 0.0}, 1e-7, 1e-14});
  }
  return buf;
}

}  // namespace array_particle

namespace array_kd_tree {

using array_particle::Particle;
constexpr size_t MAX_PARTS = 8;  // array_kd_tree.rs:14
constexpr double THETA = 0.3;    // array_kd_tree.rs:15

// `pub enum KDTree` (array_kd_tree.rs:18-34) as the flat C-ABI record; leaf_parts(i, indices) materialises the array
struct KDTree : kdnb_node {
  bool is_leaf() const { return kind == KDNB_LEAF; }
  std::array<size_t, MAX_PARTS> leaf_parts(const std::vector<size_t>& indices) const {
    std::array<size_t, MAX_PARTS> out;
    out.fill(leaf_first == KDNB_NO_INDEX ? (size_t)KDNB_NO_INDEX : 0);  // NEGS (:16) vs the 0 padding of a built leaf (:525)
    for (uint64_t k = 0; k < num_parts && k < MAX_PARTS; ++k) out[k] = indices[leaf_first + k];
    return out;
  }
};
static_assert(sizeof(KDTree) == sizeof(kdnb_node), "KDTree must stay layout-compatible with kdnb_node");

class Context {
 public:
  explicit Context(uint32_t flags = 0, int layout = KDNB_LAYOUT_PADDED, uint32_t max_parts = MAX_PARTS,
                   double theta = THETA, int device = 0) {
    kdnb_config cfg{};
    cfg.struct_size = sizeof cfg;
    cfg.device = device;
    cfg.max_parts = max_parts;
    cfg.layout = layout;
    cfg.theta = theta;
    cfg.flags = flags;
    h_ = kdnb_create(&cfg);
    if (!h_) throw std::runtime_error(std::string("kdnb_create: ") + kdnb_last_error(nullptr));
  }
  ~Context() { kdnb_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  kdnb_ctx* get() const { return h_; }
  void check(int rc, const char* what) const {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + kdnb_last_error(h_));
  }

 private:
  kdnb_ctx* h_ = nullptr;
};

// array_kd_tree.rs:45-53
inline size_t nodes_needed_for_particles(size_t num_parts) { return (size_t)kdnb_nodes_needed(num_parts, MAX_PARTS); }

// array_kd_tree.rs:55-60
inline std::vector<KDTree> allocate_node_vec(size_t num_parts) {
  KDTree def{};
  def.kind = KDNB_LEAF;
  def.leaf_first = KDNB_NO_INDEX;
  return std::vector<KDTree>(nodes_needed_for_particles(num_parts), def);
}

// array_kd_tree.rs:515-583 — whole-tree build on the GPU (cur_node must be 0, indices must cover all particles)
inline void build_tree_par4(std::vector<size_t>& indices, size_t cur_node, const std::vector<Particle>& particles,
                            std::vector<KDTree>& nodes, size_t /*thread_cnt*/ = 1) {
  if (cur_node != 0 || indices.size() != particles.size())
    throw std::runtime_error("build_tree_par4: the GPU build constructs the whole tree");
  Context c;
  c.check(kdnb_upload_particles(c.get(), particles.data(), particles.size()), "kdnb_upload_particles");
  c.check(kdnb_build_tree(c.get()), "kdnb_build_tree");
  const uint64_t need = kdnb_node_count(c.get());
  if (nodes.size() < need) throw std::runtime_error("build_tree_par4: nodes shorter than allocate_node_vec(n)");
  std::vector<uint64_t> idx(particles.size());
  uint64_t nn = 0;
  c.check(kdnb_download_tree(c.get(), nodes.data(), nodes.size(), &nn, idx.data()), "kdnb_download_tree");
  for (size_t i = 0; i < idx.size(); ++i) indices[i] = (size_t)idx[i];
}

// acc[i] = calc_accel(i, particles, tree) for every particle on a fresh tree (array_kd_tree.rs:647, :585-621)
inline std::vector<std::array<double, 3>> calc_accel_all(const std::vector<Particle>& particles) {
  Context c;
  c.check(kdnb_upload_particles(c.get(), particles.data(), particles.size()), "kdnb_upload_particles");
  c.check(kdnb_build_tree(c.get()), "kdnb_build_tree");
  c.check(kdnb_calc_accel(c.get()), "kdnb_calc_accel");
  std::vector<std::array<double, 3>> acc(particles.size());
  c.check(kdnb_download_accel(c.get(), reinterpret_cast<double*>(acc.data())), "kdnb_download_accel");  // (empty: no-op)
  return acc;
}

// array_kd_tree.rs:623-664
inline void simple_sim(std::vector<Particle>& bodies, double dt, int64_t steps) {
  Context c;
  c.check(kdnb_simple_sim_bodies(c.get(), bodies.data(), bodies.size(), dt, steps), "kdnb_simple_sim_bodies");
}

}  // namespace array_kd_tree

// ---- the Sequential crate's SIMD surface (Sequential/RustVersion/src/simd_particle.rs, simd_kd_tree.rs)
namespace simd_particle {

using Particle = kdnb_particle_simd;  // simd_particle.rs:3-8: { p: f64x4, v: f64x4, r, m }, 96 bytes, lane 3 = 0

// simd_particle.rs:10-27
inline std::vector<Particle> two_bodies() {
  std::vector<Particle> b(2);
  b[0] = Particle{{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}, 1.0, 1.0, {0.0, 0.0}};
  b[1] = Particle{{1.0, 0.0, 0.0, 0.0}, {0.0, 1.0, 0.0, 0.0}, 1e-4, 1e-20, {0.0, 0.0}};
  return b;
}

// simd_particle.rs:29-55 — n+1 particles, angles u * TAU (the array version uses 6.28); seeded splitmix64 stream
inline std::vector<Particle> circular_orbits(size_t n, uint64_t seed = 12345) {
  std::vector<Particle> buf;
  buf.reserve(n + 1);
  buf.push_back(Particle{{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}, 0.00465047, 1.0, {0.0, 0.0}});
  uint64_t st = seed;
  for (size_t i = 0; i < n; ++i) {
    const double d = 0.1 + ((double)i * 5.0 / (double)n);
    const double v = std::sqrt(1.0 / d);
    const double theta = (double)(array_particle::splitmix64(st) >> 11) * 0x1.0p-53 * 6.283185307179586;
    buf.push_back(Particle{{d * std::cos(theta), d * std::sin(theta), 0.0, 0.0},
                           {-v * std::sin(theta), v * std::cos(theta), 0.0, 0.0}, 1e-7, 1e-14, {0.0, 0.0}});
  }
  return buf;
}

}  // namespace simd_particle

namespace simd_kd_tree {

constexpr size_t MAX_PARTS = 7;  // simd_kd_tree.rs:9

// simd_kd_tree.rs:169-202 (dense `build_tree` layout, :49-138)
inline void simple_sim(std::vector<simd_particle::Particle>& bodies, double dt, int64_t steps) {
  array_kd_tree::Context c(0, KDNB_LAYOUT_DENSE, (uint32_t)MAX_PARTS);
  c.check(kdnb_simple_sim_bodies_simd(c.get(), bodies.data(), bodies.size(), dt, steps), "kdnb_simple_sim_bodies_simd");
}

}  // namespace simd_kd_tree
