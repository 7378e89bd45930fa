// sqmp_plan_multi.cpp — the planning queries of a work directory, `batch` at a time in lock step on one GPU
// (sqmp_multi.hpp), optionally from several host threads (each with its own context and stream; the voxel fields are
// built once and shared).  Same arguments and the same per-query output lines as sqmp_plan, plus:
//   sqmp_plan_multi <robot.sqmpdesc> <work_dir> gpu <iterations> <seed> [device] [first_query] [n_queries] [batch] [threads]
// There is no CPU backend in this binary: without the CUDA library / a B200 it exits with code 3.
#include <memory>
#include <thread>

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
  int threads = argc > 10 ? atoi(argv[10]) : 1;
  if (threads < 1) threads = 1;
  size_t count = a.count;
  if ((a.first + count) * 16 > a.queries.size()) count = a.queries.size() / 16 > a.first ? a.queries.size() / 16 - a.first : 0;
  if (size_t(threads) > count) threads = count ? int(count) : 1;

  std::vector<std::unique_ptr<sqmp_shim::CollisionChecker> > checkers;
  for (int t = 0; t < threads; ++t) {
    checkers.emplace_back(new sqmp_shim::CollisionChecker(argv[1], a.device));
    if (!checkers.back()->isInitialized()) return 3;
    if (t == 0) checkers[0]->setVoxelGrid(a.g);
    else if (!checkers[size_t(t)]->shareOcTreeWith(*checkers[0])) return 3;
  }
  sqmp_planner::Params P;
  P.lo[0] = a.xy_lo[0]; P.lo[1] = a.xy_lo[1]; P.hi[0] = a.xy_hi[0]; P.hi[1] = a.xy_hi[1];

  typedef sqmp_planner::MultiScheduler::Result Result;
  const size_t nt = size_t(threads);
  std::vector<std::vector<Result> > res(nt);
  std::vector<int> ok(nt, 1);
  std::vector<long> rounds(nt, 0), calls(nt, 0);
  const size_t per = (count + size_t(threads) - 1) / size_t(threads);
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t]() {
      const size_t lo = size_t(t) * per, n = lo < count ? (count - lo < per ? count - lo : per) : 0;
      sqmp_planner::GpuService service(checkers[size_t(t)].get(), true, true);
      sqmp_planner::MultiScheduler sched(&service, batch);
      // query k of the whole run keeps its seed (seed + global index) whatever the split into threads and batches
      ok[size_t(t)] = sched.run(a.queries.data() + (a.first + lo) * 16, n, a.iterations, a.seed, a.first + lo, P, res[size_t(t)]) ? 1 : 0;
      rounds[size_t(t)] = sched.rounds();
      calls[size_t(t)] = sched.library_calls();
      if (sched.profile().on && t == 0) sched.profile().print(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), sched.rounds());
    });
  }
  for (auto &th : pool) th.join();
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  long total_states = 0, total_iter = 0, total_rounds = 0, total_calls = 0;
  size_t k = 0;
  for (int t = 0; t < threads; ++t) {
    if (!ok[size_t(t)]) return 6;
    total_rounds += rounds[size_t(t)];
    total_calls += calls[size_t(t)];
    for (size_t j = 0; j < res[size_t(t)].size(); ++j, ++k) {
      if (!res[size_t(t)][j].valid_endpoints) {
        std::printf("%zu invalid_endpoint\n", a.first + k);
        continue;
      }
      const sqmp_planner::Stats &s = res[size_t(t)][j].stats;
      std::printf("%zu %d %.17g %ld %ld %ld %ld %ld %ld %ld %ld %ld %ld %llu %.17g %.17g %ld %ld\n", a.first + k, int(s.solved), s.best_cost, s.iterations,
                  s.nn_calls, s.near_calls, s.edge_batches, s.edges_checked, s.states_checked, s.rewires, s.parents_chosen,
                  s.nodes_start, s.nodes_goal, (unsigned long long)s.trace, s.best_cost_rev, s.best_cost_pri, s.informed_samples, s.solution_rewires);
      total_states += s.states_checked;
      total_iter += s.iterations;
    }
  }
  std::printf("# queries %zu seconds %.6f states %ld iterations %ld rounds %ld library_calls %ld batch %d threads %d\n", count, dt,
              total_states, total_iter, total_rounds, total_calls, batch, threads);
  return 0;
}
