// plan_oracle.cpp — TEST INFRASTRUCTURE / CPU BASELINE ONLY (like everything under oracle/).
// The planner front-end of the package (shim/sqmp_birrt.hpp, byte-for-byte the code the CUDA build runs) instantiated
// over the CPU oracle library, so that tests can compare per-iteration decisions and bench.py can time the CPU arm.
//   sqmp_plan_oracle <robot.sqmpdesc (unused)> <work_dir> oracle <iterations> <seed> [device (unused)] [first_query] [n_queries]
// work_dir additionally holds oracle_model.bin (written by tools/planner_workdir.py).
#include "sqmp_plan_driver.hpp"

#include "sqmp_oracle.h"

// CPU backend over the oracle library: the planner above it is byte-for-byte the same code as on the GPU.
class OracleBackend {
 public:
  OracleBackend(orc_model *m, unsigned flags, int tip) : m_(m), flags_(flags), tip_(tip) {}
  void sample_batch(uint64_t seed, long long first, int n, const double *lo, const double *hi, int max_tries, double *q_out) {
    orc_sample_configs(m_, tip_, seed, first, n, lo, hi, max_tries, q_out, NULL);
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
  OracleBackend be(m, 3u, tip);
  const int rc = sqmp_plan_driver::run_queries(be, a.queries, a.iterations, a.seed, a.first, a.count, a.xy_lo, a.xy_hi);
  orc_model_destroy(m);
  return rc;
}
