// sqmp_backends.hpp — the Backend of sqmp_birrt.hpp over the C ABI of libsqmp_b200.so (the product path).
#pragma once

#include <cstdio>
#include <vector>

#include "sqmp_shim.hpp"

namespace sqmp_planner {

class GpuBackend {
 public:
  GpuBackend(sqmp_shim::CollisionChecker *checker, bool check_self, bool check_map)
      : ctx_(checker->context()), flags_(sqmp_shim::CollisionChecker::flags(check_self, check_map)) {}

  void tree_reset(int slot) { ok(sqmp_tree_reset(ctx_, slot)); }
  int tree_append(int slot, const double *q, double cost) {
    int32_t id = -1;
    ok(sqmp_tree_append(ctx_, slot, q, cost, &id));
    return id;
  }
  void tree_update(int slot, int id, const double *q, double cost) { ok(sqmp_tree_update(ctx_, slot, id, q, cost)); }
  int nearest(int slot, const double *q) {
    int32_t id = 0;
    ok(sqmp_nn_nearest(ctx_, slot, q, 1, &id));
    return id;
  }
  void near(int slot, const double *q, int exclude_id, double r, int k, std::vector<int> &low, std::vector<int> &high, int &count) {
    low.assign(size_t(k), -1);
    high.assign(size_t(k), -1);
    int32_t excl = exclude_id, cnt = 0;
    ok(sqmp_nn_radius(ctx_, slot, q, &excl, 1, r, k, low.data(), high.data(), &cnt));
    count = cnt;
  }
  void sample_batch(uint64_t seed, long long first, int n, const double *lo, const double *hi, int max_tries, double *q_out) {
    ok(sqmp_sample_configs(ctx_, seed, first, n, lo, hi, max_tries, q_out, NULL));
  }
  void sample_informed_batch(uint64_t seed, long long first, int n, const double *lo, const double *hi, int max_tries,
                             const sqmp_informed &el, double *q_out) {
    ok(sqmp_sample_informed(ctx_, seed, first, n, lo, hi, max_tries, &el, q_out, NULL));
  }
  bool config_valid(const double *q) {
    uint32_t w = 1;
    if (flags_ == 0) return true;
    ok(sqmp_check_states(ctx_, q, 1, flags_, &w));
    return (w & 1u) == 0;
  }
  void check_edges(const std::vector<double> &a, const std::vector<double> &b, int num_points, std::vector<int> &first_invalid) {
    const int64_t E = int64_t(a.size() / 8);
    first_invalid.assign(size_t(E), 0);
    mask_.assign(size_t((E * (num_points + 1) + 31) / 32), 0u);
    if (flags_ == 0) {
      first_invalid.assign(size_t(E), -1);
      return;
    }
    ok(sqmp_check_edges(ctx_, a.data(), b.data(), E, num_points, flags_, mask_.data(), first_invalid.data()));
  }
  bool failed() const { return failed_; }

 private:
  void ok(int rc) {
    if (rc != SQMP_OK) {
      failed_ = true;
      std::fprintf(stderr, "libsqmp_b200: %s (%s)\n", sqmp_strerror(rc), sqmp_last_error(ctx_));
    }
  }
  sqmp_ctx *ctx_;
  unsigned flags_;
  std::vector<uint32_t> mask_;
  bool failed_ = false;
};

}  // namespace sqmp_planner
