// sqmp_plan.cpp — runs the batched BiRRT* front-end (sqmp_birrt.hpp) over planning queries on the CUDA library.
//   sqmp_plan <robot.sqmpdesc> <work_dir> gpu <iterations> <seed> [device] [first_query] [n_queries]
// work_dir holds grid_meta.txt + grid_bits.u32 (world) and queries.f64 (rows: start[8] goal[8]).
// There is no CPU backend in this binary: without the CUDA library / a B200 it exits with code 3.
#include "sqmp_backends.hpp"
#include "sqmp_plan_driver.hpp"

int main(int argc, char **argv) {
  sqmp_plan_driver::Args a;
  if (!sqmp_plan_driver::parse_args(argc, argv, a)) return 2;
  if (a.backend != "gpu") {
    std::fprintf(stderr, "unknown backend %s (this binary only has the CUDA backend)\n", a.backend.c_str());
    return 2;
  }
  sqmp_shim::CollisionChecker checker(argv[1], a.device);
  if (!checker.isInitialized()) return 3;
  checker.setVoxelGrid(a.g);
  sqmp_planner::GpuBackend be(&checker, true, true);
  return sqmp_plan_driver::run_queries(be, a.queries, a.iterations, a.seed, a.first, a.count, a.xy_lo, a.xy_hi);
}
