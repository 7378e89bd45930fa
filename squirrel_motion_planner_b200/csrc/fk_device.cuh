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
// Coefficients live in the constant bank so that every Horner step is ONE DFMA with a c[bank][offset] operand; as
// literals the compiler materialises each 64-bit constant with two UMOVs per use (3 issue slots per step).
static __constant__ double SQMP_SC64[18] = {
    0.63661977236758138243,                                   // [0] 2/pi
    1.57079632673412561417e+00, 6.07710050630396597660e-11, 2.02226624871116645580e-21,   // [1..3] pi/2 in 3 parts
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,  // [4..9] sin kernel S6..S1
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, // [10..15] cos kernel C6..C1
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02,
    -0.5, 1.0};

// fp64 sincos for joint angles: Cody-Waite reduction by pi/2 in three parts (exact for |k| < 2^20, i.e. |x| < 1e5;
// larger arguments take the library path) + the fdlibm minimax kernels on [-pi/4, pi/4].  <= ~1.5 ulp, about a
// quarter of the instructions of the general-purpose library routine (no Payne-Hanek set-up, no special cases).
// the body without the large-argument test (callers that have checked |x| < 1e5 for all their angles up front run it as
// straight-line code, which lets the compiler interleave independent chains)
__device__ __forceinline__ void sincos64_small(double x, double *sn, double *cs);
#ifdef SQMP_SINCOS_NOINLINE
// One copy per kernel instead of one per chain segment (validity_fast.cu): with every segment's sincos inlined the
// validity kernel grows past the instruction cache (no_instruction became its top stall reason).  Values travel in
// registers (double2 / float2 by value), not through the stack.
static __device__ __noinline__ double2 sincos64_call(double x) {
  double s, c;
  if (!(fabs(x) < 1.0e5)) sincos(x, &s, &c);
  else sincos64_small(x, &s, &c);
  return make_double2(s, c);
}
template <>
__device__ __forceinline__ void sincos_t<double>(double x, double *sn, double *cs) {
  const double2 v = sincos64_call(x);
  *sn = v.x;
  *cs = v.y;
}
#else
template <>
__device__ __forceinline__ void sincos_t<double>(double x, double *sn, double *cs) {
  if (!(fabs(x) < 1.0e5)) {
    sincos(x, sn, cs);
    return;
  }
  sincos64_small(x, sn, cs);
}
#endif
__device__ __forceinline__ void sincos64_small(double x, double *sn, double *cs) {
  const double *C = SQMP_SC64;
  const double kd = rint(x * C[0]);
  const int k = int(kd);
  double r = fma(-kd, C[1], x);
  r = fma(-kd, C[2], r);
  r = fma(-kd, C[3], r);
  const double z = r * r;
  double ps = fma(z, C[4], C[5]);
  ps = fma(ps, z, C[6]);
  ps = fma(ps, z, C[7]);
  ps = fma(ps, z, C[8]);
  ps = fma(ps, z, C[9]);
  const double sr = fma(ps * z, r, r);
  double pc = fma(z, C[10], C[11]);
  pc = fma(pc, z, C[12]);
  pc = fma(pc, z, C[13]);
  pc = fma(pc, z, C[14]);
  pc = fma(pc, z, C[15]);
  const double cr = fma(z * z, pc, fma(z, C[16], C[17]));
  const double s0 = (k & 1) ? cr : sr, c0 = (k & 1) ? sr : cr;
  *sn = (k & 2) ? -s0 : s0;
  *cs = ((k + 1) & 2) ? -c0 : c0;
}

// fp32 counterpart (Cody-Waite in three parts, cephes sinf/cosf kernels on [-pi/4, pi/4]): <= 2 ulp for |x| < 8e3.
#ifdef SQMP_SINCOS_NOINLINE
__device__ __forceinline__ void sincos32_body(float x, float *sn, float *cs);
static __device__ __noinline__ float2 sincos32_call(float x) {
  float s, c;
  sincos32_body(x, &s, &c);
  return make_float2(s, c);
}
template <>
__device__ __forceinline__ void sincos_t<float>(float x, float *sn, float *cs) {
  const float2 v = sincos32_call(x);
  *sn = v.x;
  *cs = v.y;
}
__device__ __forceinline__ void sincos32_body(float x, float *sn, float *cs) {
#else
template <>
__device__ __forceinline__ void sincos_t<float>(float x, float *sn, float *cs) {
#endif
  if (!(fabsf(x) < 8.0e3f)) {
    sincosf(x, sn, cs);
    return;
  }
  const float kf = rintf(x * 0.636619772f);
  const int k = int(kf);
  float r = fmaf(-kf, 1.5703125f, x);              // pi/2 = 1.5703125 + 4.837512969970703125e-4 + 7.54978995489188e-8
  r = fmaf(-kf, 4.837512969970703125e-4f, r);
  r = fmaf(-kf, 7.54978995489188e-8f, r);
  const float z = r * r;
  float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(ps, z, -1.6666654611e-1f);
  const float sr = fmaf(ps * z, r, r);
  float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(pc, z, 4.166664568298827e-2f);
  const float cr = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
  const float s0 = (k & 1) ? cr : sr, c0 = (k & 1) ? sr : cr;
  *sn = (k & 2) ? -s0 : s0;
  *cs = ((k + 1) & 2) ? -c0 : c0;
}

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
// SMALL: the caller guarantees |q| < 1e5 for rotations in fp64 (no large-argument branch in the sincos)
template <typename T, typename C, bool SMALL = false>
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
  if constexpr (SMALL && sizeof(T) == 8) {
    double sd, cd;
    sincos64_small(double(q), &sd, &cd);
    s = T(sd);
    c = T(cd);
  } else {
    sincos_t<T>(q, &s, &c);
  }
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

// ---------------------------------------------------------------------------------------------------------------
// Compile-time chain signatures.  The reference hard-codes its robot (eight named joint links, CC:216-255; 8 DOF,
// BS:6868-6870), so the structure of the chain — joint types, axis codes, which pre-frames are identity /
// translation / general — is known when this library is compiled; only the numbers come from the description at
// run time.  A kernel instantiated on a signature unrolls the chain with no interpreter (no per-segment branches,
// no jump table, constant-bank operands at fixed offsets); the host uses it only when the loaded description
// matches the signature exactly and otherwise falls back to the generic interpreter above.
// ---------------------------------------------------------------------------------------------------------------
struct SigGeneric {
  static constexpr bool generic = true;
};
struct SigRobotino {   // robotino_plan.urdf: x, y prismatic; theta about z; arm joints -y, +y, +z, -y, +z
  static constexpr bool generic = false;
  static constexpr int n = 8;
  static constexpr int first_rot = 2;   // segments 0, 1 are the prismatic base
  __host__ __device__ static constexpr int type(int s) { return s < 2 ? SQMP_JOINT_TRANS : SQMP_JOINT_ROT; }
  __host__ __device__ static constexpr int axis(int s) {
    return s == 0 ? AX_PX : s == 1 ? AX_PY : s == 2 ? AX_PZ : s == 3 ? AX_NY : s == 4 ? AX_PY : s == 5 ? AX_PZ : s == 6 ? AX_NY : AX_PZ;
  }
  __host__ __device__ static constexpr int kind(int s) { return s < 3 ? 0 : (s == 3 ? 2 : 1); }
  static constexpr int ee_kind = 2;
};
template <typename SIG>
inline bool chain_matches(const RobotDev &rb) {
  if (rb.n_seg != SIG::n || rb.ee_tip_kind != SIG::ee_kind) return false;
  for (int s = 0; s < SIG::n; ++s)
    if (rb.seg_type[s] != SIG::type(s) || rb.seg_axis_code[s] != SIG::axis(s) || rb.seg_pre_kind[s] != SIG::kind(s) ||
        rb.seg_qidx[s] != s)
      return false;
  return true;
}

template <typename T, typename SIG, int S, bool SMALL = false>
__device__ __forceinline__ void frame_advance_sig(Frame<T> &f, const RobotDev &rb, const double *q_mine, int pitch) {
  if constexpr (S < SIG::n) {
    const T qv = T(q_mine[S * pitch]);   // signature guarantees seg_qidx[s] == s
    if (sizeof(T) == 4) {
      frame_mul_const<T>(f, rb.seg_pre_f[S], SIG::kind(S));
      frame_joint_motion<T, float, SMALL>(f, SIG::type(S), SIG::axis(S), rb.seg_axis_f[S], qv);
    } else {
      frame_mul_const<T>(f, rb.seg_pre[S], SIG::kind(S));
      frame_joint_motion<T, double, SMALL>(f, SIG::type(S), SIG::axis(S), rb.seg_axis[S], qv);
    }
  }
}
template <typename T, typename SIG, int S0, int S1, bool SMALL = false>
__device__ __forceinline__ void chain_walk_sig(Frame<T> &f, const RobotDev &rb, const double *q_mine, int pitch) {
  if constexpr (S0 < S1) {
    frame_advance_sig<T, SIG, S0, SMALL>(f, rb, q_mine, pitch);
    chain_walk_sig<T, SIG, S0 + 1, S1, SMALL>(f, rb, q_mine, pitch);
  }
}

// KDL::Rotation::GetQuaternion (frames.cpp) — same branch structure so signs match the reference.  Every branch
// needs 1/sqrt(t) and sqrt(t) of one value t: both come from one rsqrt (s = t * rsqrt(t)), no division.
__device__ __forceinline__ void get_quaternion(const double *R, double &x, double &y, double &z, double &w) {
  const double trace = R[0] + R[4] + R[8];
  if (trace > 1e-12) {
    const double t = trace + 1.0, rs = rsqrt(t);
    const double s = 0.5 * rs;          // 0.5 / sqrt(t)
    w = 0.5 * (t * rs);                 // 0.25 / s
    x = (R[7] - R[5]) * s;
    y = (R[2] - R[6]) * s;
    z = (R[3] - R[1]) * s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double t = 1.0 + R[0] - R[4] - R[8], rs = rsqrt(t);
    const double is = 0.5 * rs;         // 1 / s with s = 2 sqrt(t)
    w = (R[7] - R[5]) * is;
    x = 0.5 * (t * rs);                 // 0.25 * s
    y = (R[1] + R[3]) * is;
    z = (R[2] + R[6]) * is;
  } else if (R[4] > R[8]) {
    const double t = 1.0 + R[4] - R[0] - R[8], rs = rsqrt(t);
    const double is = 0.5 * rs;
    w = (R[2] - R[6]) * is;
    x = (R[1] + R[3]) * is;
    y = 0.5 * (t * rs);
    z = (R[5] + R[7]) * is;
  } else {
    const double t = 1.0 + R[8] - R[0] - R[4], rs = rsqrt(t);
    const double is = 0.5 * rs;
    w = (R[3] - R[1]) * is;
    x = (R[2] + R[6]) * is;
    y = (R[5] + R[7]) * is;
    z = 0.5 * (t * rs);
  }
}

}  // namespace sqmp
