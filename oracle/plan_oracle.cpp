// plan_oracle.cpp — TEST INFRASTRUCTURE / CPU BASELINE ONLY (like everything under oracle/).
// The planner front-end of the package (shim/sqmp_birrt.hpp, byte-for-byte the code the CUDA build runs) instantiated
// over the CPU oracle library, so that tests can compare per-iteration decisions and bench.py can time the CPU arm.
//   sqmp_plan_oracle <robot.sqmpdesc (unused)> <work_dir> oracle <iterations> <seed> [device (unused)] [first_query] [n_queries] [batch]
// With [batch] the queries run through the lock-step scheduler of shim/sqmp_multi.hpp over the oracle.
// work_dir additionally holds oracle_model.bin (written by tools/planner_workdir.py).
#include "sqmp_multi.hpp"
#include "sqmp_plan_driver.hpp"

#include "sqmp_oracle.h"

// CPU backend over the oracle library: the planner above it is byte-for-byte the same code as on the GPU.
class OracleBackend {
 public:
  OracleBackend(orc_model *m, unsigned flags, int tip) : m_(m), flags_(flags), tip_(tip) {}
  void sample_batch(uint64_t seed, long long first, int n, const double *lo, const double *hi, int max_tries, double *q_out) {
    orc_sample_configs(m_, tip_, seed, first, n, lo, hi, max_tries, q_out, NULL);
  }
  void sample_informed_batch(uint64_t seed, long long first, int n, const double *lo, const double *hi, int max_tries,
                             const sqmp_informed &el, double *q_out) {
    static_assert(sizeof(sqmp_informed) == sizeof(orc_informed), "orc_informed mirrors sqmp_informed");
    orc_sample_informed(m_, tip_, seed, first, n, lo, hi, max_tries, reinterpret_cast<const orc_informed *>(&el), q_out, NULL);
  }
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
  int tip_;
  std::vector<double> nodes_[8], cost_[8];
};

// The "library" of the lock-step scheduler (shim/sqmp_multi.hpp) over the CPU oracle: the batched calls are loops over
// the oracle's single-tree functions.  Lets the scheduler's host logic be tested without a GPU: its lines must equal
// the sequential runner's.
class OracleService {
 public:
  OracleService(orc_model *m, unsigned flags, int tip) : m_(m), flags_(flags), tip_(tip), nodes_(2048), cost_(2048) {}
  static int max_batch() { return 1024; }
  unsigned flags() const { return flags_; }
  int tree_reset(int t) { nodes_[size_t(t)].clear(); cost_[size_t(t)].clear(); return 0; }
  int tree_append(int t, const double *q, double cost, int32_t *id) {
    nodes_[size_t(t)].insert(nodes_[size_t(t)].end(), q, q + 8);
    cost_[size_t(t)].push_back(cost);
    *id = int32_t(cost_[size_t(t)].size()) - 1;
    return 0;
  }
  int tree_update(int t, int id, const double *q, double cost) {
    std::memcpy(&nodes_[size_t(t)][size_t(id) * 8], q, 64);
    cost_[size_t(t)][size_t(id)] = cost;
    return 0;
  }
  int sample(uint64_t seed, long long first, int n, const double *lo, const double *hi, int max_tries, double *q_out) {
    orc_sample_configs(m_, tip_, seed, first, n, lo, hi, max_tries, q_out, NULL);
    return 0;
  }
  int sample_informed(uint64_t seed, long long first, int n, const double *lo, const double *hi, int max_tries, const sqmp_informed &el,
                      double *q_out) {
    orc_sample_informed(m_, tip_, seed, first, n, lo, hi, max_tries, reinterpret_cast<const orc_informed *>(&el), q_out, NULL);
    return 0;
  }
  int nearest_multi(const int32_t *tree, const double *q, int64_t n, int32_t *ids) {
    for (int64_t i = 0; i < n; ++i)
      ids[i] = orc_nearest(nodes_[size_t(tree[i])].data(), int64_t(cost_[size_t(tree[i])].size()), q + i * 8);
    return 0;
  }
  int radius_multi(const int32_t *tree, const double *q, const int32_t *excl, int64_t n, double r, int k, int32_t *low,
                   int32_t *high, int32_t *cnt) {
    for (int64_t i = 0; i < n; ++i)
      orc_near_topk_batch(nodes_[size_t(tree[i])].data(), cost_[size_t(tree[i])].data(), int64_t(cost_[size_t(tree[i])].size()),
                          q + i * 8, excl + i, 1, r, k, low + i * k, high + i * k, cnt + i, 1);
    return 0;
  }
  int check_edges(const double *a, const double *b, int64_t E, int num_points, uint32_t *mask, int32_t *first) {
    orc_check_edges(m_, a, b, E, num_points, flags_, mask, first, 1);
    return 0;
  }
  int check_states(const double *q, int64_t n, uint32_t *mask) {
    for (int64_t i = 0; i < n; ++i) {
      if ((i & 31) == 0) mask[i >> 5] = 0u;
      if (orc_is_in_collision(m_, q + i * 8, (flags_ & 1u) != 0, (flags_ & 2u) != 0)) mask[i >> 5] |= 1u << (i & 31);
    }
    return 0;
  }
  void report(int) const {}

 private:
  orc_model *m_;
  unsigned flags_;
  int tip_;
  std::vector<std::vector<double> > nodes_, cost_;
};

// oracle_model.bin (written by tests / bench): int32 L, n_spheres; then the arrays orc_model_create takes, in order;
// then int32 tip = link index of the FK chain tip (hand_wrist_link, SRDF:4-6)
static orc_model *load_oracle_model(const std::string &path, int *tip) {
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
  int32_t tip32 = -1;
  rd(&tip32, 4);
  if (!f || tip32 < 0 || tip32 >= L) return NULL;
  *tip = tip32;
  return orc_model_create(L, parent.data(), jt.data(), qidx.data(), jo.data(), ja.data(), tp.data(), ft.data(), shaped.data(),
                          pair.data(), n, sph_link.data(), sc.data(), sr.data());
}

int main(int argc, char **argv) {
  sqmp_plan_driver::Args a;
  if (!sqmp_plan_driver::parse_args(argc, argv, a)) return 2;
  if (a.backend != "oracle") {
    std::fprintf(stderr, "unknown backend %s (this binary only has the CPU oracle backend)\n", a.backend.c_str());
    return 2;
  }
  int tip = -1;
  orc_model *m = load_oracle_model(a.dir + "oracle_model.bin", &tip);
  if (!m) return 4;
  const double origin[3] = {a.g.origin[0], a.g.origin[1], a.g.origin[2] - 0.02};   // CC:87, as the shim does
  orc_set_grid(m, a.g.bits.data(), a.g.gx, a.g.gy, a.g.gz, origin, a.g.res);
  int rc = 0;
  if (argc > 9) {   // [batch]: the lock-step scheduler over the oracle (same output lines as the sequential runner)
    const char *fl = getenv("SQMP_ORACLE_FLAGS");   // profiling aid: 0 = no collision checks (host logic only)
    OracleService svc(m, fl ? unsigned(atoi(fl)) : 3u, tip);
    sqmp_planner::MultiSchedulerT<OracleService> sched(&svc, atoi(argv[9]));
    sqmp_planner::Params P;
    P.lo[0] = a.xy_lo[0]; P.lo[1] = a.xy_lo[1]; P.hi[0] = a.xy_hi[0]; P.hi[1] = a.xy_hi[1];
    size_t count = a.count;
    if ((a.first + count) * 16 > a.queries.size()) count = a.queries.size() / 16 > a.first ? a.queries.size() / 16 - a.first : 0;
    std::vector<sqmp_planner::MultiSchedulerT<OracleService>::Result> res;
    if (!sched.run(a.queries.data() + a.first * 16, count, a.iterations, a.seed, a.first, P, res)) rc = 6;
    for (size_t k = 0; k < res.size() && rc == 0; ++k) {
      if (!res[k].valid_endpoints) {
        std::printf("%zu invalid_endpoint\n", a.first + k);
        continue;
      }
      const sqmp_planner::Stats &s = res[k].stats;
      std::printf("%zu %d %.17g %ld %ld %ld %ld %ld %ld %ld %ld %ld %ld %llu %.17g %.17g %ld %ld\n", a.first + k, int(s.solved), s.best_cost,
                  s.iterations, s.nn_calls, s.near_calls, s.edge_batches, s.edges_checked, s.states_checked, s.rewires,
                  s.parents_chosen, s.nodes_start, s.nodes_goal, (unsigned long long)s.trace, s.best_cost_rev, s.best_cost_pri, s.informed_samples, s.solution_rewires);
    }
    std::printf("# queries %zu rounds %ld service_calls %ld\n", count, sched.rounds(), sched.library_calls());
  } else {
    OracleBackend be(m, 3u, tip);
    rc = sqmp_plan_driver::run_queries(be, a.queries, a.iterations, a.seed, a.first, a.count, a.xy_lo, a.xy_hi);
  }
  orc_model_destroy(m);
  return rc;
}
