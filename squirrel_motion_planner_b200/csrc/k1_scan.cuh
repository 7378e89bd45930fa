// k1_scan.cuh — the exact stage shared by the validity kernels: sphere vs occupied-voxel boxes by bit-row scans, and
// the per-warp queue that balances those scans over the lanes.  (Reference semantics replaced: fcl::collide of a link
// shape against the octree, CC:554-592 — here "some occupied voxel AABB within r of the sphere centre".)
#pragma once
#include "fk_device.cuh"

namespace sqmp {

// ------------------------------------------------------------------------------------------------
// exact sphere vs occupied-voxel-boxes test, in GRID UNITS: voxel (i,j,k) is the unit cube [i,i+1]x[j,j+1]x[k,k+1],
// the sphere has centre u and radius rg = r / res.  true <=> dist(u, some occupied unit cube) <= rg.
// Per (j,k) row the admissible x-interval is closed-form, so a row costs one (rarely two) word loads.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool exact_scan(const GridDev &g, double ux, double uy, double uz, double rg) {
  const double r2 = rg * rg;
  const int k0 = max(__double2int_rd(uz - rg), 0), k1 = min(__double2int_rd(uz + rg), g.gz - 1);
  const int j0 = max(__double2int_rd(uy - rg), 0), j1 = min(__double2int_rd(uy + rg), g.gy - 1);
  double kd = double(k0);
  for (int k = k0; k <= k1; ++k, kd += 1.0) {
    const double dz = fmax(fmax(kd - uz, uz - (kd + 1.0)), 0.0);
    const double remz = fma(-dz, dz, r2);
    if (remz < 0.0) continue;
    const uint32_t *slab = g.occ + size_t(k) * g.gy * g.row_words;
    double jd = double(j0);
    for (int j = j0; j <= j1; ++j, jd += 1.0) {
      const double dy = fmax(fmax(jd - uy, uy - (jd + 1.0)), 0.0);
      const double rem = fma(-dy, dy, remz);
      if (rem < 0.0) continue;
      const double w = sqrt(rem);  // half-width of the admissible x interval
      // cube i is within reach  <=>  i + 1 >= ux - w  and  i <= ux + w
      int i0 = max(__double2int_ru(ux - w) - 1, 0);
      const int i1 = min(__double2int_rd(ux + w), g.gx - 1);
      const uint32_t *row = slab + size_t(j) * g.row_words;
      while (i0 <= i1) {  // at most 2*ceil(rg)+2 wide: one word, two when it straddles a word boundary
        const int sh = i0 & 31;
        const int n = min(32 - sh, i1 - i0 + 1);
        const uint32_t bits = (row[i0 >> 5] >> sh) & (n >= 32 ? ~0u : ((1u << n) - 1u));
        if (bits) return true;
        i0 += n;
      }
    }
  }
  return false;
}

// `count` consecutive rows of the (j,k) scan box of one sphere, starting at row (j0 + jj, k), j fastest.
__device__ __forceinline__ bool scan_rows(const GridDev &g, double ux, double uy, double uz, double rg, int j0, int nj,
                                          int k, int jj, int count) {
  const double r2 = rg * rg;
  for (int n = 0; n < count; ++n) {
    const double kd = double(k), jd = double(j0 + jj);
    const double dz = fmax(fmax(kd - uz, uz - (kd + 1.0)), 0.0);
    const double dy = fmax(fmax(jd - uy, uy - (jd + 1.0)), 0.0);
    const double rem = fma(-dy, dy, fma(-dz, dz, r2));
    if (rem >= 0.0) {
      const double w = sqrt(rem);
      int i0 = max(__double2int_ru(ux - w) - 1, 0);
      const int i1 = min(__double2int_rd(ux + w), g.gx - 1);
      const uint32_t *row = g.occ + (size_t(k) * g.gy + size_t(j0 + jj)) * g.row_words;
      while (i0 <= i1) {
        const int sh = i0 & 31;
        const int m = min(32 - sh, i1 - i0 + 1);
        const uint32_t bits = (row[i0 >> 5] >> sh) & (m >= 32 ? ~0u : ((1u << m) - 1u));
        if (bits) return true;
        i0 += m;
      }
    }
    if (++jj == nj) { jj = 0; ++k; }
  }
  return false;
}

// Same rows, decided by a CONSERVATIVE fp32 filter in coordinates relative to the sphere's own cell (all magnitudes
// <= ~15, one float ulp <= 1e-6 cell): per row the admissible x-interval is computed from an inflated radius and an
// inflated half-width, so it is a superset of the exact one; a row whose superset window holds no occupied bit is
// free — the common case, because the spheres that reach this stage are mostly free-but-close.  Only a row with a
// bit in its window is re-decided by the fp64 code above, so the verdict is the fp64 verdict, bit for bit.
// The NR rows of a chunk are handled together: all windows first, then all word loads back to back (one memory
// round trip per chunk instead of one per row), then the tests.  ~30 fp32/integer instructions per row.
template <int NR>
__device__ __forceinline__ bool scan_rows_filtered(const GridDev &g, double ux, double uy, double uz, double rg, int j0,
                                                   int nj, int k, int jj, int count) {
  const int cx = __double2int_rd(ux), cy = __double2int_rd(uy), cz = __double2int_rd(uz);
  const float fx = float(ux - double(cx)), fy = float(uy - double(cy)), fz = float(uz - double(cz));   // in [0,1]
  const float fy1 = fy - 1.0f, fz1 = fz - 1.0f;
  const float rf = float(rg);
  const float r2 = fmaf(rf * rf, 1.00002f, 2e-5f);   // >= rg^2 + every rounding error below (rg <= 13 cells)
  const unsigned rw = unsigned(g.row_words);
  uint32_t w0[NR], w1[NR], msk[NR];
  int shv[NR];
  int kq = k, jq = jj;
  float oyf = float(j0 + jj - cy), ozf = float(k - cz);   // row offsets from the own cell, stepped in float
  unsigned ridx = (unsigned(k) * unsigned(g.gy) + unsigned(j0 + jj)) * rw;   // occ words < 2^32 (host)
#pragma unroll
  for (int n = 0; n < NR; ++n) {
    w0[n] = w1[n] = msk[n] = 0u;
    shv[n] = 0;
    if (n < count) {
      // gap between f in [0,1] and the unit cube at offset o: max(o - f, (f - 1) - o, 0)
      const float gy = fmaxf(fmaxf(oyf - fy, fy1 - oyf), 0.0f);
      const float gz = fmaxf(fmaxf(ozf - fz, fz1 - ozf), 0.0f);
      const float rem = fmaf(-gy, gy, fmaf(-gz, gz, r2));
      if (rem >= 0.0f) {
        const float w = fmaf(rem, rsqrtf(fmaxf(rem, 1e-12f)), 2e-5f);   // >= sqrt(rem) + slack
        const int i0 = max(cx + __float2int_ru(fx - w) - 1, 0);
        const int i1 = min(cx + __float2int_rd(fx + w), g.gx - 1);
        if (i0 <= i1) {   // window of <= 2 * 14 + 1 bits: two adjacent words (the row has a pad word), one funnel shift
          const uint32_t *p = g.occ + (ridx + unsigned(i0 >> 5));
#ifdef SQMP_BOUNDS_CHECK   // debug builds only: the row word (and its right neighbour) must lie inside the occupancy rows
          if (kq < 0 || kq >= g.gz || j0 + jq < 0 || j0 + jq >= g.gy || (i0 >> 5) + 1 >= g.row_words) __trap();
#endif
          const int sh = i0 & 31, width = i1 - i0 + 1;
          w0[n] = p[0];
          if (sh + width > 32) w1[n] = p[1];
          shv[n] = sh;
          msk[n] = width >= 32 ? ~0u : ((1u << width) - 1u);
        }
      }
      ridx += rw;
      oyf += 1.0f;
      if (++jq == nj) {   // next slab of the scan box
        jq = 0;
        ++kq;
        ozf += 1.0f;
        oyf = float(j0 - cy);
        ridx = (unsigned(kq) * unsigned(g.gy) + unsigned(j0)) * rw;
      }
    }
  }
  unsigned any = 0;
#pragma unroll
  for (int n = 0; n < NR; ++n)
    if (__funnelshift_r(w0[n], w1[n], shv[n]) & msk[n]) any |= 1u << n;
  while (any) {   // rare: decide the flagged rows exactly
    const int n = __ffs(any) - 1;
    any &= any - 1;
    const int lin = jj + n, dk = lin / nj;
    if (scan_rows(g, ux, uy, uz, rg, j0, nj, k + dk, lin - dk * nj, 1)) return true;
  }
  return false;
}

// ------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------
constexpr int K1_THREADS = 128;
constexpr int K1_QCAP = 64;       // deferred exact-scan items per warp (drained 32 at a time)
#ifndef K1_CHUNK_DEF
#define K1_CHUNK_DEF 3
#endif
constexpr int K1_CHUNK = K1_CHUNK_DEF;   // rows of an exact scan handed to one lane at a time (their loads are issued together); 3 measured best (2, 4, 6, 8 tried)       // rows of an exact scan handed to one lane at a time (their loads are issued together)
constexpr int K1_BATCH = 4;       // spheres whose field gathers are in flight together
constexpr unsigned FULL = 0xffffffffu;

struct WarpQueue {   // shared memory, one per warp
  double x[K1_QCAP], y[K1_QCAP], z[K1_QCAP], r[K1_QCAP];
  int info[K1_QCAP];   // sphere index << 8 | owner lane
};

// round-to-nearest integer of x without F2I: the integer lands in the low mantissa bits of (x + 1.5 * 2^mant)
__device__ __forceinline__ int rn_int(double x) { return __double2loint(x + 6755399441055744.0); }       // |x| < 2^31
__device__ __forceinline__ int rn_int(float x) { return __float_as_int(x + 12582912.0f) - 0x4B400000; }   // |x| < 2^22

// gap between coordinate f in [0,1] of the centre's own cell and the unit cube at cell offset o
__device__ __forceinline__ float axis_gap_off(int o, float f) {
  return o > 0 ? float(o) - f : (o < 0 ? f - float(o) - 1.0f : 0.0f);
}

// Decide the queued items [base, base+n); returns the verdicts as a mask over the owning lanes.  Items differ by
// two orders of magnitude in work (a 1-cell finger sphere scans ~9 rows, the 5-cell head sphere ~150), so the
// unit handed to a lane is not an item but a chunk of <= K1_CHUNK rows: lane l describes item l, a warp prefix
// sum numbers the chunks, and every lane finds the item of its chunk by a 5-step shuffle binary search.
// One copy in the binary (called from the two drain sites of K1): keeps the hot loop inside the instruction cache.
static __device__ __noinline__ unsigned drain_queue(const GridDev &gd, const WarpQueue *wq, int base, int n, unsigned known) {
  const int lane = threadIdx.x & 31;
  int nj = 1, j0 = 0, k0 = 0, R = 0, c = 0;
  if (lane < n) {
    const int owner = wq->info[base + lane] & 31;
    if (!((known >> owner) & 1u)) {
      const double uy = wq->y[base + lane], uz = wq->z[base + lane], rg = wq->r[base + lane];
      k0 = max(__double2int_rd(uz - rg), 0);
      j0 = max(__double2int_rd(uy - rg), 0);
      const int k1 = min(__double2int_rd(uz + rg), gd.gz - 1), j1 = min(__double2int_rd(uy + rg), gd.gy - 1);
      nj = max(j1 - j0 + 1, 1);
      R = (j1 >= j0 && k1 >= k0) ? nj * (k1 - k0 + 1) : 0;
      c = (R + K1_CHUNK - 1) / K1_CHUNK;
    }
  }
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(FULL, incl, 31);
  unsigned hm = 0;
  for (int t0 = 0; t0 < total; t0 += 32) {
    const int t = t0 + lane;
    int m = 0;  // first item whose inclusive chunk count exceeds t
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const int v = __shfl_sync(FULL, incl, m + step - 1);
      if (v <= t) m += step;
    }
    const int ex_m = __shfl_sync(FULL, incl - c, m), nj_m = __shfl_sync(FULL, nj, m), j0_m = __shfl_sync(FULL, j0, m),
              k0_m = __shfl_sync(FULL, k0, m), R_m = __shfl_sync(FULL, R, m);
    if (t < total) {
      const double ux = wq->x[base + m], uy = wq->y[base + m], uz = wq->z[base + m], rg = wq->r[base + m];
      const int r0 = (t - ex_m) * K1_CHUNK, r1 = min(r0 + K1_CHUNK, R_m);
      const int kk = r0 / nj_m;
#ifdef SQMP_SCAN_UNFILTERED   // A/B builds only (tools/build_variant.sh): the fp64 scan without the fp32 row filter
      if (scan_rows(gd, ux, uy, uz, rg, j0_m, nj_m, k0_m + kk, r0 - kk * nj_m, r1 - r0))
#else
      if (scan_rows_filtered<K1_CHUNK>(gd, ux, uy, uz, rg, j0_m, nj_m, k0_m + kk, r0 - kk * nj_m, r1 - r0))
#endif
        hm |= 1u << (wq->info[base + m] & 31);
    }
  }
  return __reduce_or_sync(FULL, hm);
}

// ------------------------------------------------------------------------------------------------
// the state of one thread: either given (row sl of q_in), or interpolated from the end points of edge sl / (P + 1)
// with the reference's exact operation order (BS:4461-4462, 4478; no FMA); then parked in shared memory, prismatic
// values converted to grid units.  Returns false for a non-finite or absurd joint value: the kernels report such a state
// "in collision" (fail closed, the spirit of CC:106-107; the oracle restates the same rule).
// ------------------------------------------------------------------------------------------------
template <bool EDGES>
__device__ __forceinline__ void k1_load_state(const double *__restrict__ q_in, const double *__restrict__ q_end, int num_points,
                                              int64_t sl, double (&q)[SQMP_DOF]) {
  if (EDGES) {
    const int P1 = num_points + 1;
    const int64_t e = sl / P1;
    const int kk = int(sl - e * P1);
    const double4 *a4 = reinterpret_cast<const double4 *>(q_in + e * SQMP_DOF);
    const double4 *b4 = reinterpret_cast<const double4 *>(q_end + e * SQMP_DOF);
    const double4 a0 = a4[0], a1 = a4[1], b0 = b4[0], b1 = b4[1];
    const double st[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const double en[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const double np = double(num_points), fk = double(kk);
#pragma unroll
    for (int j = 0; j < SQMP_DOF; ++j)
      q[j] = __dadd_rn(st[j], __dmul_rn(fk, __ddiv_rn(__dsub_rn(en[j], st[j]), np)));
  } else {
    const double4 *p4 = reinterpret_cast<const double4 *>(q_in + sl * SQMP_DOF);
    const double4 v0 = p4[0], v1 = p4[1];
    q[0] = v0.x; q[1] = v0.y; q[2] = v0.z; q[3] = v0.w;
    q[4] = v1.x; q[5] = v1.y; q[6] = v1.z; q[7] = v1.w;
  }
}
__device__ __forceinline__ bool k1_park_state(const RobotDev &rb, const double (&q)[SQMP_DOF], double *q_mine, int pitch) {
  bool sane = true;
#pragma unroll
  for (int j = 0; j < SQMP_DOF; ++j) {
    const double v = ((rb.trans_mask >> j) & 1) ? q[j] * rb.len_scale : q[j];
    sane = sane && (fabs(v) < 1.0e9) && (fabs(q[j]) < 1.0e7);   // the fail-closed rule of the boundary: |q_j| < 1e7
    q_mine[j * pitch] = v;
  }
  return sane;
}

}  // namespace sqmp
