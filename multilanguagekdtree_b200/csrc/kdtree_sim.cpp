// kdtree-sim — the CLI of Parallel/RustVersion/src/main.rs on the B200 path.
//   kdtree-sim --number N [--steps S]        (clap: -n/--number required, -s/--steps default 1; main.rs:8-18)
//   kdtree-sim <steps> <n> [threads]         (the positional form of the reference's other CLIs, e.g.
//                                             Parallel/CppVersion/kdtree-sim.cpp:14-16, driven by Parallel/*/benchmark.sh;
//                                             `threads` is accepted and ignored: the work runs on the GPU)
// dt = 1e-3 (main.rs:23).  Prints the elapsed seconds of circular_orbits + simple_sim like main.rs:25-31.
// Extras (not in the reference): --seed K, --verbose (per-stage device milliseconds on stderr).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kdnb.hpp"

static void usage() {
  std::fprintf(stderr, "Usage: kdtree-sim --number <NUMBER> [--steps <STEPS>] [--seed <SEED>] [--verbose]\n"
                       "       kdtree-sim <STEPS> <NUMBER> [THREADS]\n");
}

int main(int argc, char** argv) {
  long long number = -1, steps = 1;
  unsigned long long seed = 12345;
  bool verbose = false;
  int npos = 0;
  for (int i = 1; i < argc; ++i) {
    auto is = [&](const char* a, const char* b) { return !std::strcmp(argv[i], a) || !std::strcmp(argv[i], b); };
    if (is("-n", "--number") && i + 1 < argc) number = std::atoll(argv[++i]);
    else if (is("-s", "--steps") && i + 1 < argc) steps = std::atoll(argv[++i]);
    else if (is("--seed", "--seed") && i + 1 < argc) seed = std::strtoull(argv[++i], nullptr, 10);
    else if (is("-v", "--verbose")) verbose = true;
    else if (is("-h", "--help")) { usage(); return 0; }
    else if (argv[i][0] >= '0' && argv[i][0] <= '9' && npos < 3) {  // positional: steps, n, threads
      const long long v = std::atoll(argv[i]);
      if (npos == 0) steps = v;
      else if (npos == 1) number = v;
      ++npos;
    }
    else { std::fprintf(stderr, "error: unexpected argument '%s'\n", argv[i]); usage(); return 2; }
  }
  if (npos == 1) {
    std::fprintf(stderr, "Specify a number of steps and a number of particles.\n");
    return 1;
  }
  if (number < 0) {
    std::fprintf(stderr, "error: the following required arguments were not provided:\n  --number <NUMBER>\n");
    usage();
    return 2;
  }
  const double dt = 1e-3;
  try {
    const auto start = std::chrono::steady_clock::now();
    auto bodies = array_particle::circular_orbits((size_t)number, seed);
    if (!verbose) {
      array_kd_tree::simple_sim(bodies, dt, steps);
    } else {
      array_kd_tree::Context c(KDNB_FLAG_PROFILE);
      c.check(kdnb_simple_sim_bodies(c.get(), bodies.data(), bodies.size(), dt, steps), "kdnb_simple_sim_bodies");
      double ms[KDNB_STAGE_COUNT];
      uint64_t n = 0;
      c.check(kdnb_stage_ms(c.get(), ms, &n), "kdnb_stage_ms");
      std::fprintf(stderr, "steps=%llu build=%.3f ms walk=%.3f ms kick=%.3f ms exchange=%.3f ms launches=%llu\n",
                   (unsigned long long)n, ms[KDNB_STAGE_BUILD], ms[KDNB_STAGE_WALK], ms[KDNB_STAGE_KICK],
                   ms[KDNB_STAGE_EXCHANGE], (unsigned long long)kdnb_launch_count(c.get()));
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    std::printf("%.9g\n", secs);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "kdtree-sim: %s\n", e.what());
    return 1;
  }
  return 0;
}
