// fk_device.cuh — per-thread forward kinematics of the planning chain (device code).
//
// Reference semantics: CollisionChecker::updateTransforms (collision_checker.hpp:519-539) and
// KDL::ChainFkSolverPos_recursive as used by KDLRobotModel::compute_FK (kdl_kuka_model.cpp:278-305).
// Not a port: the 39-link walk is folded by the host into <= 8 segments (robot_desc.py), each segment is
// "fixed pre-frame, then a joint motion about a (usually signed unit) axis", and a thread keeps only the
// running frame in registers.  T is double (default; verdicts then agree with the fp64 oracle to ~1e-13 m)
// or float (arm chain only; tolerance budget in DESIGN.md).
#pragma once
#include "sqmp_common.cuh"

namespace sqmp {

template <typename T>
struct Frame {
  T r[9];  // row-major
  T p[3];
};

template <typename T>
__device__ __forceinline__ void sincos_t(T x, T *s, T *c);
// fp64 sincos for joint angles: Cody-Waite reduction by pi/2 in three parts (exact for |k| < 2^20, i.e. |x| < 1e5;
// larger arguments take the library path) + the fdlibm minimax kernels on [-pi/4, pi/4].  <= ~1.5 ulp, about a
// quarter of the instructions of the general-purpose library routine (no Payne-Hanek set-up, no special cases).
template <>
__device__ __forceinline__ void sincos_t<double>(double x, double *sn, double *cs) {
  if (!(fabs(x) < 1.0e5)) {
    sincos(x, sn, cs);
    return;
  }
  const double kd = rint(x * 0.63661977236758138243);  // 2/pi
  const int k = int(kd);
  double r = fma(-kd, 1.57079632673412561417e+00, x);  // pi/2 split: 33 + 33 + 53 bits (fdlibm pio2_1, pio2_2, pio2_3)
  r = fma(-kd, 6.07710050630396597660e-11, r);
  r = fma(-kd, 2.02226624871116645580e-21, r);
  const double z = r * r;
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  const double sr = fma(ps * z, r, r);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
  const double s0 = (k & 1) ? cr : sr, c0 = (k & 1) ? sr : cr;
  *sn = (k & 2) ? -s0 : s0;
  *cs = ((k + 1) & 2) ? -c0 : c0;
}
template <>
__device__ __forceinline__ void sincos_t<float>(float x, float *s, float *c) { sincosf(x, s, c); }

template <typename T>
__device__ __forceinline__ void frame_set_translation(Frame<T> &f, T x, T y, T z) {
  f.r[0] = T(1); f.r[1] = T(0); f.r[2] = T(0);
  f.r[3] = T(0); f.r[4] = T(1); f.r[5] = T(0);
  f.r[6] = T(0); f.r[7] = T(0); f.r[8] = T(1);
  f.p[0] = x; f.p[1] = y; f.p[2] = z;
}

// f <- f * Frame(R, t) with the constant 12-vector c (R rows, t); kind: 0 identity, 1 translation, 2 general
template <typename T, typename C>
__device__ __forceinline__ void frame_mul_const(Frame<T> &f, const C *c, int kind) {
  if (kind == 0) return;
  const T tx = T(c[9]), ty = T(c[10]), tz = T(c[11]);
#pragma unroll
  for (int i = 0; i < 3; ++i) f.p[i] = fma(f.r[3 * i + 0], tx, fma(f.r[3 * i + 1], ty, fma(f.r[3 * i + 2], tz, f.p[i])));
  if (kind == 2) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const T a = f.r[3 * i + 0], b = f.r[3 * i + 1], d = f.r[3 * i + 2];
      f.r[3 * i + 0] = fma(a, T(c[0]), fma(b, T(c[3]), d * T(c[6])));
      f.r[3 * i + 1] = fma(a, T(c[1]), fma(b, T(c[4]), d * T(c[7])));
      f.r[3 * i + 2] = fma(a, T(c[2]), fma(b, T(c[5]), d * T(c[8])));
    }
  }
}

// columns (ca, cb) <- (c*ca + s*cb, -s*ca + c*cb): right-multiplication by a rotation about the third axis
template <typename T>
__device__ __forceinline__ void rot_cols(Frame<T> &f, int ca, int cb, T s, T c) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const T a = f.r[3 * i + ca], b = f.r[3 * i + cb];
    f.r[3 * i + ca] = fma(c, a, s * b);
    f.r[3 * i + cb] = fma(c, b, -s * a);
  }
}

// f <- f * Motion(axis, q):  KDL::Joint::pose(q) without its origin (already folded into the pre-frame)
template <typename T, typename C>
__device__ __forceinline__ void frame_joint_motion(Frame<T> &f, int type, int axis_code, const C *axis, T q) {
  if (type == SQMP_JOINT_TRANS) {
    if (axis_code == AX_GENERAL) {
      const T ax = T(axis[0]) * q, ay = T(axis[1]) * q, az = T(axis[2]) * q;
#pragma unroll
      for (int i = 0; i < 3; ++i) f.p[i] = fma(f.r[3 * i + 0], ax, fma(f.r[3 * i + 1], ay, fma(f.r[3 * i + 2], az, f.p[i])));
    } else {
      const int col = (axis_code - 1) >> 1;
      const T qq = ((axis_code - 1) & 1) ? -q : q;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const T rc = col == 0 ? f.r[3 * i + 0] : (col == 1 ? f.r[3 * i + 1] : f.r[3 * i + 2]);
        f.p[i] = fma(rc, qq, f.p[i]);
      }
    }
    return;
  }
  T s, c;
  sincos_t<T>(q, &s, &c);
  switch (axis_code) {
    case AX_PX: rot_cols(f, 1, 2, s, c); break;
    case AX_NX: rot_cols(f, 1, 2, -s, c); break;
    case AX_PY: rot_cols(f, 2, 0, s, c); break;
    case AX_NY: rot_cols(f, 2, 0, -s, c); break;
    case AX_PZ: rot_cols(f, 0, 1, s, c); break;
    case AX_NZ: rot_cols(f, 0, 1, -s, c); break;
    default: {
      // Rodrigues (KDL::Rotation::Rot2) then a full 3x3 product
      const T vx = T(axis[0]), vy = T(axis[1]), vz = T(axis[2]), vt = T(1) - c;
      T m[9];
      m[0] = fma(vt * vx, vx, c);       m[1] = fma(vt * vx, vy, -vz * s); m[2] = fma(vt * vx, vz, vy * s);
      m[3] = fma(vt * vx, vy, vz * s);  m[4] = fma(vt * vy, vy, c);       m[5] = fma(vt * vy, vz, -vx * s);
      m[6] = fma(vt * vx, vz, -vy * s); m[7] = fma(vt * vy, vz, vx * s);  m[8] = fma(vt * vz, vz, c);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const T a = f.r[3 * i + 0], b = f.r[3 * i + 1], d = f.r[3 * i + 2];
        f.r[3 * i + 0] = fma(a, m[0], fma(b, m[3], d * m[6]));
        f.r[3 * i + 1] = fma(a, m[1], fma(b, m[4], d * m[7]));
        f.r[3 * i + 2] = fma(a, m[2], fma(b, m[5], d * m[8]));
      }
    }
  }
}

// One chain segment: f <- f * pre[s] * Motion(axis[s], q[qidx[s]])
template <typename T>
__device__ __forceinline__ void frame_advance(Frame<T> &f, const RobotDev &rb, int s, const double (&q)[SQMP_DOF]) {
  // q index is warp-uniform; select without dynamic register indexing
  const int qi = rb.seg_qidx[s];
  double qv = q[0];
#pragma unroll
  for (int j = 1; j < SQMP_DOF; ++j) qv = (qi == j) ? q[j] : qv;
  if (sizeof(T) == 4) {
    frame_mul_const<T>(f, rb.seg_pre_f[s], rb.seg_pre_kind[s]);
    frame_joint_motion<T>(f, rb.seg_type[s], rb.seg_axis_code[s], rb.seg_axis_f[s], T(qv));
  } else {
    frame_mul_const<T>(f, rb.seg_pre[s], rb.seg_pre_kind[s]);
    frame_joint_motion<T>(f, rb.seg_type[s], rb.seg_axis_code[s], rb.seg_axis[s], T(qv));
  }
}

// Same, with the joint vector parked in shared memory (SoA, q_s[j * pitch + thread]): one LDS instead of a
// select chain, and 16 registers less per thread for the kernel that keeps many other values live (K1).
template <typename T>
__device__ __forceinline__ void frame_advance_smem(Frame<T> &f, const RobotDev &rb, int s, const double *q_mine, int pitch) {
  const double qv = q_mine[rb.seg_qidx[s] * pitch];
  if (sizeof(T) == 4) {
    frame_mul_const<T>(f, rb.seg_pre_f[s], rb.seg_pre_kind[s]);
    frame_joint_motion<T>(f, rb.seg_type[s], rb.seg_axis_code[s], rb.seg_axis_f[s], T(qv));
  } else {
    frame_mul_const<T>(f, rb.seg_pre[s], rb.seg_pre_kind[s]);
    frame_joint_motion<T>(f, rb.seg_type[s], rb.seg_axis_code[s], rb.seg_axis[s], T(qv));
  }
}

// KDL::Rotation::GetQuaternion (frames.cpp) — same branch structure so signs match the reference.
__device__ __forceinline__ void get_quaternion(const double *R, double &x, double &y, double &z, double &w) {
  const double trace = R[0] + R[4] + R[8];
  if (trace > 1e-12) {
    const double s = 0.5 / sqrt(trace + 1.0);
    w = 0.25 / s;
    x = (R[7] - R[5]) * s;
    y = (R[2] - R[6]) * s;
    z = (R[3] - R[1]) * s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double s = 2.0 * sqrt(1.0 + R[0] - R[4] - R[8]);
    w = (R[7] - R[5]) / s;
    x = 0.25 * s;
    y = (R[1] + R[3]) / s;
    z = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    const double s = 2.0 * sqrt(1.0 + R[4] - R[0] - R[8]);
    w = (R[2] - R[6]) / s;
    x = (R[1] + R[3]) / s;
    y = 0.25 * s;
    z = (R[5] + R[7]) / s;
  } else {
    const double s = 2.0 * sqrt(1.0 + R[8] - R[0] - R[4]);
    w = (R[3] - R[1]) / s;
    x = (R[2] + R[6]) / s;
    y = (R[5] + R[7]) / s;
    z = 0.25 * s;
  }
}

}  // namespace sqmp
