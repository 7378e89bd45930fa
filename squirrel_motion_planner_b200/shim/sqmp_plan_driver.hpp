// sqmp_plan_driver.hpp — shared pieces of the planner driver binaries: raw-file reader, the work-directory world and
// the query loop.  Backend-agnostic (template): shim/sqmp_plan.cpp instantiates it over the CUDA library; the test /
// CPU-baseline binary that instantiates it over the CPU checker lives outside the package, under oracle/.
// One line per query: index solved best_cost iterations nn near batches edges states rewires parents nodes_start nodes_goal trace
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "sqmp_birrt.hpp"
#include "sqmp_shim.hpp"

namespace sqmp_plan_driver {

template <class T>
inline std::vector<T> read_raw(const std::string &path) {
  std::ifstream f(path.c_str(), std::ios::binary | std::ios::ate);
  std::vector<T> v;
  if (!f) return v;
  const std::streamsize n = f.tellg();
  f.seekg(0);
  v.resize(size_t(n) / sizeof(T));
  f.read(reinterpret_cast<char *>(v.data()), n);
  return v;
}

// work_dir/grid_meta.txt (gx gy gz ox oy oz res) + grid_bits.u32
inline sqmp_shim::VoxelGrid read_world(const std::string &dir) {
  sqmp_shim::VoxelGrid g;
  std::ifstream meta((dir + "grid_meta.txt").c_str());
  meta >> g.gx >> g.gy >> g.gz >> g.origin[0] >> g.origin[1] >> g.origin[2] >> g.res;
  g.bits = read_raw<uint32_t>(dir + "grid_bits.u32");
  return g;
}

template <class Backend>
inline int run_queries(Backend &be, const std::vector<double> &queries, long iterations, uint64_t seed, size_t first, size_t count,
                       const double *xy_lo, const double *xy_hi) {
  const auto t0 = std::chrono::steady_clock::now();
  long total_states = 0, total_iter = 0;
  for (size_t k = first; k < first + count && (k + 1) * 16 <= queries.size(); ++k) {
    sqmp_planner::Params P;
    P.seed = seed + k;
    P.lo[0] = xy_lo[0]; P.lo[1] = xy_lo[1]; P.hi[0] = xy_hi[0]; P.hi[1] = xy_hi[1];
    sqmp_planner::BiRRTStar<Backend> planner(be, P);
    if (!planner.init(&queries[k * 16], &queries[k * 16 + 8])) {
      std::printf("%zu invalid_endpoint\n", k);
      continue;
    }
    const sqmp_planner::Stats &s = planner.run(iterations);
    std::printf("%zu %d %.17g %ld %ld %ld %ld %ld %ld %ld %ld %ld %ld %llu %.17g %.17g %ld %ld\n", k, int(s.solved), s.best_cost, s.iterations,
                s.nn_calls, s.near_calls, s.edge_batches, s.edges_checked, s.states_checked, s.rewires, s.parents_chosen,
                s.nodes_start, s.nodes_goal, (unsigned long long)s.trace, s.best_cost_rev, s.best_cost_pri, s.informed_samples, s.solution_rewires);
    total_states += s.states_checked;
    total_iter += s.iterations;
    if (be.failed()) return 6;
  }
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("# queries %zu seconds %.6f states %ld iterations %ld\n", count, dt, total_states, total_iter);
  return 0;
}


struct Args {
  std::string dir, backend;
  long iterations;
  uint64_t seed;
  int device;
  size_t first, count;
  std::vector<double> queries;
  sqmp_shim::VoxelGrid g;
  double xy_lo[2], xy_hi[2];
};

// argv: robot.sqmpdesc work_dir backend iterations seed [device] [first_query] [n_queries]
inline bool parse_args(int argc, char **argv, Args &a) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s robot.sqmpdesc work_dir backend iterations seed [device] [first_query] [n_queries]\n", argv[0]);
    return false;
  }
  a.dir = std::string(argv[2]) + "/";
  a.backend = argv[3];
  a.iterations = atol(argv[4]);
  a.seed = strtoull(argv[5], NULL, 10);
  a.device = argc > 6 ? atoi(argv[6]) : 0;
  a.queries = read_raw<double>(a.dir + "queries.f64");
  a.first = argc > 7 ? size_t(atol(argv[7])) : 0;
  a.count = argc > 8 ? size_t(atol(argv[8])) : a.queries.size() / 16;
  a.g = read_world(a.dir);
  a.xy_lo[0] = a.g.origin[0] + 1.0; a.xy_lo[1] = a.g.origin[1] + 1.0;
  a.xy_hi[0] = a.g.origin[0] + a.g.gx * a.g.res - 1.0; a.xy_hi[1] = a.g.origin[1] + a.g.gy * a.g.res - 1.0;
  return true;
}

}  // namespace sqmp_plan_driver
