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
// level above cannot rule it out, so a state far from self contact costs ~one test per block.  `blocks` and `as` may
// point to shared-memory copies of SelfDev's tables (validity_fast.cu stages them: the loops below are chains of
// dependent, warp-uniform loads, so their latency is the cost).
__device__ __forceinline__ bool self_collision(const SelfBlock *blocks, int n_blocks, const SelfA *as, const SelfHPair *hpairs,
                                               const float *f_mine, int pitch) {
  int pfa = -1, pfb = -1;
  float FA[12], FB[12], lx = 0.0f, ly = 0.0f, lz = 0.0f, ax, ay, az;
  for (int bi = 0; bi < n_blocks; ++bi) {
    const SelfBlock B = blocks[bi];
    if (B.fa != pfa || B.fb != pfb) {   // blocks come sorted by group pair (warp-uniform): frames and B bound once per pair
      pfa = B.fa;
      pfb = B.fb;
      self_load_frame(f_mine, pitch, B.fa, FA);
      self_load_frame(f_mine, pitch, B.fb, FB);
      float bx, by, bz;
      self_xform(FB, B.cb[0], B.cb[1], B.cb[2], bx, by, bz);   // bounding sphere of the pair's B side, shared by its blocks
      // ... expressed in A's frame (R_A^T (b - t_A)): A-side clusters and A-spheres are then tested as the constants
      // they are, 7 instructions each instead of a 9-FMA transform
      const float dx = bx - FA[9], dy = by - FA[10], dz = bz - FA[11];
      lx = fmaf(FA[0], dx, fmaf(FA[3], dy, FA[6] * dz));
      ly = fmaf(FA[1], dx, fmaf(FA[4], dy, FA[7] * dz));
      lz = fmaf(FA[2], dx, fmaf(FA[5], dy, FA[8] * dz));
    }
    if (self_dist2(B.ca[0], B.ca[1], B.ca[2], lx, ly, lz) > B.thr2) continue;   // this A-side cluster is out of reach
    // four A-spheres, then four partners, per step: the loops are chains of dependent warp-uniform loads and short
    // arithmetic, so their cost is latency; batching gives the scheduler four independent chains (tables are padded
    // by four records, out-of-range records are masked)
    for (int ai = B.a_begin; ai < B.a_end; ai += 4) {
      SelfA A[4];
      unsigned pass = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) A[k] = as[ai + k];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ai + k < B.a_end && self_dist2(A[k].cx, A[k].cy, A[k].cz, lx, ly, lz) <= A[k].thr2) pass |= 1u << k;
      while (pass) {
        const int k = __ffs(pass) - 1;
        pass &= pass - 1;
        const SelfA Ak = k == 0 ? A[0] : (k == 1 ? A[1] : (k == 2 ? A[2] : A[3]));
        self_xform(FA, Ak.cx, Ak.cy, Ak.cz, ax, ay, az);
        // this A-sphere in B's frame (R_B^T (a - t_B)): its partners are then constants too
        const float ex = ax - FB[9], ey = ay - FB[10], ez = az - FB[11];
        const float mx = fmaf(FB[0], ex, fmaf(FB[3], ey, FB[6] * ez));
        const float my = fmaf(FB[1], ex, fmaf(FB[4], ey, FB[7] * ez));
        const float mz = fmaf(FB[2], ex, fmaf(FB[5], ey, FB[8] * ez));
        for (int p = Ak.p_begin; p < Ak.p_end; p += 4) {
          bool h = false;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const SelfHPair pr = hpairs[p + j];
            h |= (p + j < Ak.p_end) && self_dist2(mx, my, mz, pr.bx, pr.by, pr.bz) <= pr.rsum2;
          }
          if (h) return true;
        }
      }
    }
  }
  return false;
}

}  // namespace sqmp
