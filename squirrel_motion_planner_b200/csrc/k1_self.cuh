// k1_self.cuh — self collision over the enabled cross-group sphere pairs (CollisionChecker::checkSelfCollision,
// CC:541-552; pair table CC:353-393), hierarchically (SelfDev in sqmp_common.cuh).  Shared by the validity kernels.
// The caller has parked the frame of every rigid group that owns a pair in shared memory: f_mine[(slot * 12 + k) *
// pitch], k = 0..8 rotation (row-major), 9..11 translation, fp32, all translations relative to one common point.
#pragma once
#include "sqmp_common.cuh"

namespace sqmp {

__device__ __forceinline__ void self_load_frame(const float *f_mine, int pitch, int fi, float (&F)[12]) {
#pragma unroll
  for (int k = 0; k < 12; ++k) F[k] = f_mine[(fi * 12 + k) * pitch];
}
__device__ __forceinline__ void self_xform(const float (&F)[12], float cx, float cy, float cz, float &x, float &y, float &z) {
  x = fmaf(F[0], cx, fmaf(F[1], cy, fmaf(F[2], cz, F[9])));
  y = fmaf(F[3], cx, fmaf(F[4], cy, fmaf(F[5], cz, F[10])));
  z = fmaf(F[6], cx, fmaf(F[7], cy, fmaf(F[8], cz, F[11])));
}
__device__ __forceinline__ float self_dist2(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  return fmaf(dx, dx, fmaf(dy, dy, dz * dz));
}

// true <=> some enabled pair of spheres of this state overlaps.  A level of the hierarchy is entered only if the
// level above cannot rule it out, so a state far from self contact costs ~one test per block.
__device__ __forceinline__ bool self_collision(const SelfDev &self, const RobotDev &rb, const float *f_mine, int pitch) {
  for (int bi = 0; bi < self.n_blocks; ++bi) {
    const SelfBlock B = self.blocks[bi];
    float FA[12], FB[12], ax, ay, az, bx, by, bz;
    self_load_frame(f_mine, pitch, B.fa, FA);
    self_load_frame(f_mine, pitch, B.fb, FB);
    self_xform(FA, B.ca[0], B.ca[1], B.ca[2], ax, ay, az);
    self_xform(FB, B.cb[0], B.cb[1], B.cb[2], bx, by, bz);   // bounding sphere of the block's B side
    if (self_dist2(ax, ay, az, bx, by, bz) > B.thr2) continue;
    for (int ai = B.a_begin; ai < B.a_end; ++ai) {
      const SelfA A = self.as[ai];
      const float4 ca = rb.sph_f[A.a];
      self_xform(FA, ca.x, ca.y, ca.z, ax, ay, az);
      if (self_dist2(ax, ay, az, bx, by, bz) > A.thr2) continue;
      for (int p = A.p_begin; p < A.p_end; ++p) {
        const SelfPair pr = self.pairs[p];
        const float4 cb = rb.sph_f[pr.b];
        float px, py, pz;
        self_xform(FB, cb.x, cb.y, cb.z, px, py, pz);
        if (self_dist2(ax, ay, az, px, py, pz) <= pr.rsum2) return true;
      }
    }
  }
  return false;
}

}  // namespace sqmp
