// sqmp_plan_multi.cpp — the planning queries of a work directory, `batch` at a time in lock step on one GPU
// (sqmp_multi.hpp).  Same arguments and the same per-query output lines as sqmp_plan, plus the batch size:
//   sqmp_plan_multi <robot.sqmpdesc> <work_dir> gpu <iterations> <seed> [device] [first_query] [n_queries] [batch]
// There is no CPU backend in this binary: without the CUDA library / a B200 it exits with code 3.
#include "sqmp_multi.hpp"
#include "sqmp_plan_driver.hpp"

int main(int argc, char **argv) {
  sqmp_plan_driver::Args a;
  if (!sqmp_plan_driver::parse_args(argc, argv, a)) return 2;
  if (a.backend != "gpu") {
    std::fprintf(stderr, "unknown backend %s (this binary only has the CUDA backend)\n", a.backend.c_str());
    return 2;
  }
  const int batch = argc > 9 ? atoi(argv[9]) : 128;
  sqmp_shim::CollisionChecker checker(argv[1], a.device);
  if (!checker.isInitialized()) return 3;
  checker.setVoxelGrid(a.g);
  size_t count = a.count;
  if ((a.first + count) * 16 > a.queries.size()) count = a.queries.size() / 16 > a.first ? a.queries.size() / 16 - a.first : 0;
  sqmp_planner::Params P;
  P.lo[0] = a.xy_lo[0]; P.lo[1] = a.xy_lo[1]; P.hi[0] = a.xy_hi[0]; P.hi[1] = a.xy_hi[1];
  sqmp_planner::GpuService service(&checker, true, true);
  sqmp_planner::MultiScheduler sched(&service, batch);
  std::vector<sqmp_planner::MultiScheduler::Result> res;
  const auto t0 = std::chrono::steady_clock::now();
  const bool ok = sched.run(a.queries.data() + a.first * 16, count, a.iterations, a.seed, a.first, P, res);
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (!ok) return 6;
  long total_states = 0, total_iter = 0;
  for (size_t k = 0; k < res.size(); ++k) {
    if (!res[k].valid_endpoints) {
      std::printf("%zu invalid_endpoint\n", a.first + k);
      continue;
    }
    const sqmp_planner::Stats &s = res[k].stats;
    std::printf("%zu %d %.17g %ld %ld %ld %ld %ld %ld %ld %ld %ld %ld %llu\n", a.first + k, int(s.solved), s.best_cost, s.iterations,
                s.nn_calls, s.near_calls, s.edge_batches, s.edges_checked, s.states_checked, s.rewires, s.parents_chosen,
                s.nodes_start, s.nodes_goal, (unsigned long long)s.trace);
    total_states += s.states_checked;
    total_iter += s.iterations;
  }
  std::printf("# queries %zu seconds %.6f states %ld iterations %ld rounds %ld library_calls %ld batch %d\n", count, dt,
              total_states, total_iter, sched.rounds(), sched.library_calls(), batch);
  return 0;
}
