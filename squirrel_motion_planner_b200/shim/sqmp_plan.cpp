// sqmp_plan.cpp — runs the batched BiRRT* front-end (sqmp_birrt.hpp) over planning queries.
//   sqmp_plan <robot.sqmpdesc> <work_dir> <backend: gpu|oracle> <iterations> <seed> [device] [first_query] [n_queries]
// work_dir holds grid_meta.txt + grid_bits.u32 (world), queries.f64 (rows: start[8] goal[8]) and, for the oracle
// backend (TEST / CPU-BASELINE USE ONLY, compiled in only with -DSQMP_WITH_ORACLE), oracle_model.bin.
// One line per query: index solved best_cost iterations nn near batches edges states rewires parents nodes_start nodes_goal trace
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "sqmp_backends.hpp"
#include "sqmp_birrt.hpp"

#ifdef SQMP_WITH_ORACLE
#include "sqmp_oracle.h"
#endif

template <class T>
static std::vector<T> read_raw(const std::string &path) {
  std::ifstream f(path.c_str(), std::ios::binary | std::ios::ate);
  std::vector<T> v;
  if (!f) return v;
  const std::streamsize n = f.tellg();
  f.seekg(0);
  v.resize(size_t(n) / sizeof(T));
  f.read(reinterpret_cast<char *>(v.data()), n);
  return v;
}

#ifdef SQMP_WITH_ORACLE
// CPU backend over the oracle library: the planner above it is byte-for-byte the same code as on the GPU.
class OracleBackend {
 public:
  OracleBackend(orc_model *m, unsigned flags) : m_(m), flags_(flags) {}
  void tree_reset(int slot) { nodes_[slot].clear(); cost_[slot].clear(); }
  int tree_append(int slot, const double *q, double cost) {
    nodes_[slot].insert(nodes_[slot].end(), q, q + 8);
    cost_[slot].push_back(cost);
    return int(cost_[slot].size()) - 1;
  }
  void tree_update(int slot, int id, const double *q, double cost) {
    std::memcpy(&nodes_[slot][size_t(id) * 8], q, 64);
    cost_[slot][size_t(id)] = cost;
  }
  int nearest(int slot, const double *q) { return orc_nearest(nodes_[slot].data(), int64_t(cost_[slot].size()), q); }
  void near(int slot, const double *q, int exclude_id, double r, int k, std::vector<int> &low, std::vector<int> &high, int &count) {
    low.assign(size_t(k), -1);
    high.assign(size_t(k), -1);
    int32_t excl = exclude_id, cnt = 0;
    orc_near_topk_batch(nodes_[slot].data(), cost_[slot].data(), int64_t(cost_[slot].size()), q, &excl, 1, r, k, low.data(),
                        high.data(), &cnt, 1);
    count = cnt;
  }
  bool config_valid(const double *q) { return !orc_is_in_collision(m_, q, (flags_ & 1u) != 0, (flags_ & 2u) != 0); }
  void check_edges(const std::vector<double> &a, const std::vector<double> &b, int num_points, std::vector<int> &first_invalid) {
    const int64_t E = int64_t(a.size() / 8);
    first_invalid.assign(size_t(E), 0);
    std::vector<uint32_t> mask(size_t((E * (num_points + 1) + 31) / 32), 0u);
    orc_check_edges(m_, a.data(), b.data(), E, num_points, flags_, mask.data(), first_invalid.data(), 1);
  }
  bool failed() const { return false; }

 private:
  orc_model *m_;
  unsigned flags_;
  std::vector<double> nodes_[8], cost_[8];
};

// oracle_model.bin (written by tests / bench): int32 L, n_spheres; then the arrays orc_model_create takes, in order
static orc_model *load_oracle_model(const std::string &path) {
  std::ifstream f(path.c_str(), std::ios::binary);
  if (!f) return NULL;
  int32_t L = 0, n = 0;
  f.read(reinterpret_cast<char *>(&L), 4);
  f.read(reinterpret_cast<char *>(&n), 4);
  std::vector<int> parent(L), jt(L), qidx(L), shaped(L), sph_link(n);
  std::vector<double> jo(size_t(L) * 3), ja(size_t(L) * 3), tp(size_t(L) * 12), ft(size_t(L) * 12), sc(size_t(n) * 3), sr(n);
  std::vector<unsigned char> pair(size_t(L) * L);
  auto rd = [&](void *p, size_t bytes) { f.read(reinterpret_cast<char *>(p), std::streamsize(bytes)); };
  rd(parent.data(), 4 * L); rd(jt.data(), 4 * L); rd(qidx.data(), 4 * L); rd(jo.data(), 24 * L); rd(ja.data(), 24 * L);
  rd(tp.data(), 96 * L); rd(ft.data(), 96 * L); rd(shaped.data(), 4 * L); rd(pair.data(), size_t(L) * L);
  rd(sph_link.data(), 4 * n); rd(sc.data(), 24 * n); rd(sr.data(), 8 * n);
  if (!f) return NULL;
  return orc_model_create(L, parent.data(), jt.data(), qidx.data(), jo.data(), ja.data(), tp.data(), ft.data(), shaped.data(),
                          pair.data(), n, sph_link.data(), sc.data(), sr.data());
}
#endif

template <class Backend>
static int run_queries(Backend &be, const std::vector<double> &queries, long iterations, uint64_t seed, size_t first, size_t count,
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
    std::printf("%zu %d %.17g %ld %ld %ld %ld %ld %ld %ld %ld %ld %ld %llu\n", k, int(s.solved), s.best_cost, s.iterations,
                s.nn_calls, s.near_calls, s.edge_batches, s.edges_checked, s.states_checked, s.rewires, s.parents_chosen,
                s.nodes_start, s.nodes_goal, (unsigned long long)s.trace);
    total_states += s.states_checked;
    total_iter += s.iterations;
    if (be.failed()) return 6;
  }
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("# queries %zu seconds %.6f states %ld iterations %ld\n", count, dt, total_states, total_iter);
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s robot.sqmpdesc work_dir gpu|oracle iterations seed [device] [first_query] [n_queries]\n", argv[0]);
    return 2;
  }
  const std::string dir = std::string(argv[2]) + "/", backend = argv[3];
  const long iterations = atol(argv[4]);
  const uint64_t seed = strtoull(argv[5], NULL, 10);
  const int device = argc > 6 ? atoi(argv[6]) : 0;
  const std::vector<double> queries = read_raw<double>(dir + "queries.f64");
  const size_t first = argc > 7 ? size_t(atol(argv[7])) : 0, count = argc > 8 ? size_t(atol(argv[8])) : queries.size() / 16;
  sqmp_shim::VoxelGrid g;
  {
    std::ifstream meta((dir + "grid_meta.txt").c_str());
    meta >> g.gx >> g.gy >> g.gz >> g.origin[0] >> g.origin[1] >> g.origin[2] >> g.res;
    g.bits = read_raw<uint32_t>(dir + "grid_bits.u32");
  }
  const double xy_lo[2] = {g.origin[0] + 1.0, g.origin[1] + 1.0};
  const double xy_hi[2] = {g.origin[0] + g.gx * g.res - 1.0, g.origin[1] + g.gy * g.res - 1.0};
  if (backend == "gpu") {
    sqmp_shim::CollisionChecker checker(argv[1], device);
    if (!checker.isInitialized()) return 3;
    checker.setVoxelGrid(g);
    sqmp_planner::GpuBackend be(&checker, true, true);
    return run_queries(be, queries, iterations, seed, first, count, xy_lo, xy_hi);
  }
#ifdef SQMP_WITH_ORACLE
  if (backend == "oracle") {
    orc_model *m = load_oracle_model(dir + "oracle_model.bin");
    if (!m) return 4;
    const double origin[3] = {g.origin[0], g.origin[1], g.origin[2] - 0.02};   // CC:87, as the shim does
    orc_set_grid(m, g.bits.data(), g.gx, g.gy, g.gz, origin, g.res);
    OracleBackend be(m, 3u);
    const int rc = run_queries(be, queries, iterations, seed, first, count, xy_lo, xy_hi);
    orc_model_destroy(m);
    return rc;
  }
#endif
  std::fprintf(stderr, "unknown backend %s\n", backend.c_str());
  return 2;
}
