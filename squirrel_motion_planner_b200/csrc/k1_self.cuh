// k1_self.cuh — self collision over the enabled link pairs whose links hang on different rigid groups
// (CollisionChecker::checkSelfCollision, CC:541-552; pair table CC:353-393), by nested bounds (SelfDev in
// sqmp_common.cuh).  Shared by the validity kernels.
// The caller has parked the frame of every rigid group that takes part in shared memory: f_mine[(slot * 12 + k) *
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
// squared distance from a point to an axis-aligned box
__device__ __forceinline__ float self_box_dist2(float px, float py, float pz, const float (&lo)[3], const float (&hi)[3]) {
  const float dx = fmaxf(fmaxf(lo[0] - px, px - hi[0]), 0.0f), dy = fmaxf(fmaxf(lo[1] - py, py - hi[1]), 0.0f),
              dz = fmaxf(fmaxf(lo[2] - pz, pz - hi[2]), 0.0f);
  return fmaf(dx, dx, fmaf(dy, dy, dz * dz));
}
// T = FA^-1 * FB: a point given in B's frame expressed in A's frame
__device__ __forceinline__ void self_relative(const float (&FA)[12], const float (&FB)[12], float (&T)[12]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) T[3 * i + j] = fmaf(FA[i], FB[j], fmaf(FA[3 + i], FB[3 + j], FA[6 + i] * FB[6 + j]));
  const float dx = FB[9] - FA[9], dy = FB[10] - FA[10], dz = FB[11] - FA[11];
#pragma unroll
  for (int i = 0; i < 3; ++i) T[9 + i] = fmaf(FA[i], dx, fmaf(FA[3 + i], dy, FA[6 + i] * dz));
}

// true <=> some sphere of link a overlaps some sphere of link b, for an enabled link pair (a, b) on different rigid
// groups.  A level of the hierarchy is entered only if the level above cannot rule it out: group pair -> (link of A,
// group B) -> (cluster of A, group B) -> (cluster of A, link of B) -> cluster pair -> (cluster of A, leaf of B) -> leaf
// pair: the A side (the base: twelve links, ~45 clusters) is narrowed down against the small bound of the whole B
// group before the B side's many links are looked at.  A-side bounds are axis-aligned BOXES in A's frame (they are
// constants there, and flat shapes like the shell box or the drum fill a box far better than a sphere), B-side bounds
// are spheres (they travel through the relative transform as points).  Everything is expressed in A's
// frame, so the A side is tested as the constants it is.  LINK_HITS (diagnostics): no early exit, every colliding link
// pair is recorded in link_pair_hit[64 * 64].
template <bool LINK_HITS>
__device__ __forceinline__ bool self_collision(const SelfDev &sd, const float *f_mine, int pitch, uint8_t *link_pair_hit,
                                               int gp0 = 0, int gp_step = 1) {
  bool any = false;
  for (int gi = gp0; gi < sd.n_gpairs; gi += gp_step) {
    const SelfGroupPair G = sd.gpairs[gi];
    float FA[12], FB[12], T[12];
    self_load_frame(f_mine, pitch, G.fa, FA);
    self_load_frame(f_mine, pitch, G.fb, FB);
    self_relative(FA, FB, T);
    float gx, gy, gz;   // bound of the B-side links, in A's frame
    self_xform(T, G.bx, G.by, G.bz, gx, gy, gz);
    {
      const float rs = G.RA + G.RB;
      if (self_dist2(gx, gy, gz, G.ax, G.ay, G.az) > rs * rs) continue;
    }
    for (int al = G.al_begin; al < G.al_end; ++al) {
      const SelfALink AL = sd.alinks[al];
      if (self_box_dist2(gx, gy, gz, AL.lo, AL.hi) > G.RB * G.RB) continue;   // this A link is out of reach of all of B
      const SelfLink LA = sd.links[AL.a];
      for (int ca = LA.cl_begin; ca < LA.cl_end; ++ca) {
        const SelfCluster CA = sd.clusters[ca];
        if (self_box_dist2(gx, gy, gz, CA.lo, CA.hi) > G.RB * G.RB) continue;   // ... and so is this cluster of it
        for (int lp = AL.lp_begin; lp < AL.lp_end; ++lp) {
          const SelfLink LB = sd.links[sd.lpairs[lp]];
          if (LINK_HITS && link_pair_hit[min(LA.link, LB.link) * SQMP_MAX_LINKS + max(LA.link, LB.link)]) continue;
          float px, py, pz;
          self_xform(T, LB.x, LB.y, LB.z, px, py, pz);
          if (self_box_dist2(px, py, pz, CA.lo, CA.hi) > LB.R * LB.R) continue;
          bool pair_hit = false;
          for (int cb = LB.cl_begin; cb < LB.cl_end && !pair_hit; ++cb) {
            const SelfCluster CB = sd.clusters[cb];
            float qx, qy, qz;
            self_xform(T, CB.x, CB.y, CB.z, qx, qy, qz);
            if (self_box_dist2(qx, qy, qz, CA.lo, CA.hi) > CB.R * CB.R) continue;
            for (int b = CB.leaf_begin; b < CB.leaf_end && !pair_hit; ++b) {
              const SelfLeaf Bl = sd.leaves[b];
              float bx, by, bz;
              self_xform(T, Bl.x, Bl.y, Bl.z, bx, by, bz);
              if (self_box_dist2(bx, by, bz, CA.lo, CA.hi) > Bl.r * Bl.r) continue;
              for (int a = CA.leaf_begin; a < CA.leaf_end; ++a) {
                const SelfLeaf Al = sd.leaves[a];
                const float rs = Al.r + Bl.r;
                if (self_dist2(bx, by, bz, Al.x, Al.y, Al.z) <= rs * rs) { pair_hit = true; break; }
              }
            }
          }
          if (pair_hit) {
            if (!LINK_HITS) return true;
            any = true;
            link_pair_hit[min(LA.link, LB.link) * SQMP_MAX_LINKS + max(LA.link, LB.link)] = 1;
          }
        }
      }
    }
  }
  return any;
}

}  // namespace sqmp
