// sqmp_informed.hpp — host side of the informed sampler: BiRRTstarPlanner::jointConfigEllipseInitialization
// (birrt_star.cpp:3472-3604) and the diagonal "L" of sampleJointConfigfromEllipse_JntArray (birrt_star.cpp:3709-3752).
// The samples themselves are drawn on the device (sqmp_sample_informed, include/sqmp_b200.h).
//
// Per joint class (revolute / prismatic) the reference builds centre = (start + goal) / 2, a1 = (goal - start) / c_min,
// M = a1 e1^T, U S V^T = Eigen::JacobiSVD(M), C = U diag(1, .., 1, det U det V^T) V — a ROTATION with first column a1.
// Eigen is not part of this tree, and the remaining columns of C are not determined by M (its null space has n-1
// dimensions: they fall out of Eigen's sweep order), so C is the Householder reflection e1 -> a1 with its last column
// negated: orthogonal, det +1, first column a1.  For two joints that rotation is unique.
// (squirrel_motion_planner_b200/informed.py is the same construction for the Python harness.)
#pragma once

#include <cmath>
#include <cstring>

#include "sqmp_b200.h"

namespace sqmp_shim {

class InformedEllipse {
 public:
  // start / goal: 8 joint values; is_revolute[j] != 0: joint j belongs to the revolute class
  void init(const double *start, const double *goal, const int *is_revolute) {
    n_[0] = n_[1] = 0;
    for (int j = 0; j < SQMP_DOF; ++j) {
      const int cls = is_revolute[j] ? 0 : 1;
      idx_[cls][n_[cls]++] = j;
    }
    for (int cls = 0; cls < 2; ++cls) {
      const int n = n_[cls];
      double d[SQMP_DOF], a1[SQMP_DOF], sum = 0.0;
      for (int a = 0; a < n; ++a) {
        const int j = idx_[cls][a];
        center_[cls][a] = (start[j] + goal[j]) / 2.0;               // BS:3507 / 3519
        d[a] = goal[j] - start[j];
        sum += d[a] * d[a];                                          // cost_h of the start node (DH:160-217)
      }
      c_min_[cls] = std::sqrt(sum);
      for (int a = 0; a < n; ++a) a1[a] = c_min_[cls] > 0.0 ? d[a] / c_min_[cls] : (a == 0 ? 1.0 : 0.0);   // BS:3508 / 3520
      rotation_to(a1, n, C_[cls]);
    }
  }
  // pack for sqmp_sample_informed: L from the best solution's class costs (BS:3709-3752); env = base joints' sampling box
  void fill(sqmp_informed &el, double c_best_rev, double c_best_pri, const double *lo, const double *hi) const {
    std::memset(&el, 0, sizeof(el));
    el.n_rev = n_[0];
    el.n_pri = n_[1];
    for (int cls = 0; cls < 2; ++cls) {
      const int n = n_[cls];
      int32_t *idx = cls == 0 ? el.rev_idx : el.pri_idx;
      double *Cm = cls == 0 ? el.c_rev : el.c_pri, *cen = cls == 0 ? el.center_rev : el.center_pri, *L = cls == 0 ? el.l_rev : el.l_pri;
      const double cb = cls == 0 ? c_best_rev : c_best_pri;
      for (int a = 0; a < n; ++a) {
        idx[a] = idx_[cls][a];
        cen[a] = center_[cls][a];
        for (int b = 0; b < n; ++b) Cm[a * n + b] = C_[cls][a * n + b];
      }
      const double rest = cb * cb - c_min_[cls] * c_min_[cls];
      L[0] = cb / 2.0;
      L[1] = std::sqrt(rest > 0.0 ? rest : 0.0) / 2.0;              // (clamped: rounding can make the difference negative)
    }
    el.env_x[0] = lo[0]; el.env_x[1] = hi[0];                       // m_env_size_x / m_env_size_y (BS:3810-3816)
    el.env_y[0] = lo[1]; el.env_y[1] = hi[1];
  }
  double c_min(int cls) const { return c_min_[cls]; }

  // rotation (orthogonal, det +1) whose first column is the unit vector a1; row-major n x n
  static void rotation_to(const double *a1, int n, double *C) {
    for (int i = 0; i < n * n; ++i) C[i] = 0.0;
    for (int i = 0; i < n; ++i) C[i * n + i] = 1.0;
    if (n < 2) return;
    double v[SQMP_DOF], vv = 0.0;
    for (int i = 0; i < n; ++i) {
      v[i] = a1[i] - (i == 0 ? 1.0 : 0.0);
      vv += v[i] * v[i];
    }
    if (vv == 0.0) return;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) C[i * n + j] -= (2.0 / vv) * (v[i] * v[j]);
    for (int i = 0; i < n; ++i) C[i * n + (n - 1)] = -C[i * n + (n - 1)];
  }

 private:
  int n_[2] = {0, 0};
  int idx_[2][SQMP_DOF];
  double center_[2][SQMP_DOF], c_min_[2] = {0.0, 0.0}, C_[2][SQMP_DOF * SQMP_DOF];
};

}  // namespace sqmp_shim
