// validity_fast.cu — K1F: the state-validity kernel the library uses whenever its preconditions hold (sqmp_create /
// sqmp_set_voxel_grid decide; otherwise the generic K1 of validity.cu runs).  Same predicate, same verdicts.
//
// Reference path replaced (file:line in /root/reference): CollisionChecker::isInCollision CC:104-121, updateTransforms
// CC:519-539, checkMapCollision CC:554-592, checkSelfCollision CC:541-552; BiRRTstarPlanner::interpolateConfigurations
// BS:4443-4490, isEdgeValid BS:6864-6895.
//
// The collision model is a CONSERVATIVE sphere cover of the reference's link shapes (several hundred spheres,
// sphere_model.py).  Testing every sphere of every state is what the generic K1 does; here a state pays for the
// spheres near an obstacle only:
//   * per rigid group a two-level bound hierarchy: root bound -> cluster bounds (<= 8 spheres of one link) -> spheres.
//     Bounds and spheres are "nodes" decided by the same stage 1 — one byte gather from a lower-bound field,
//     `field > thr` proves every point of the node farther than its radius from every occupied voxel — and a level
//     is entered only by the lanes whose node above is not provably free (warp-uniform loops over the union);
//   * the base group only translates in x, y and turns about z, so the z of its nodes never changes: they are tested
//     against 2-D SLAB fields (squared column distance within the layers the node can touch).  That is what makes a
//     cover of the low shapes affordable: the floor slab 4 cm under the shell is simply not in their slabs, while the
//     3-D cell-granular field would leave every low sphere undecided;
//   * the chain is walked in fp64 (identical frames to K1); per rigid group the frame is rounded ONCE to fp32 in
//     coordinates relative to the base's own cell, so |fp32 centre - fp64 centre| <= err (host-computed bound, ~5e-5
//     cells; capi.cu: refresh_grid_robot).  Stage 1 runs two nodes per packed fp32x2 instruction (9 FFMA2 + 3 FADD2
//     per pair), 2 IMAD + 1 add per node for the cell offset, one byte gather from the halo-padded field — no bounds
//     test, because the fields cover every cell a node centre can reach while the base is inside the grid.
//     "field > thr" with thr = floor((rg + err)^2) proves the fp64 centre free as well.  A 2-D node is the same
//     instruction stream with the third frame row replaced by (0 0 1 | 0) and its slab index as z;
//   * stage 2 (spheres stage 1 cannot prove free): the feature-field upper bound in fp32 against (rg - err)^2 proves a
//     hit for the fp64 centre as well;
//   * stage 3 (neither bound decides): the fp64 centre is formed from the fp64 frame exactly as K1 forms it and queued
//     for the same exact bit-row scan (k1_scan.cuh).
// The node tables travel in the kernel parameters (constant bank, warp-uniform indices); the per-sphere data of the
// rare stages comes from global memory.
// States whose base lies outside the grid never use the fp32 path: every gather of theirs reads one zero byte, which
// keeps every node open, and their spheres go through the reach test and the exact scan only.  A state with a
// non-finite or absurd joint value is reported in collision (fail closed).
#include <cstdlib>
#include <cstring>

#ifndef SQMP_KF_INLINE_SINCOS
// one sincos body per kernel instead of one per chain segment: 4000 -> 2900 instructions.  ncu had no_instructions (instruction
// cache misses) as the top stall of the fully inlined kernel; out of line it is 4 % faster (0.378 -> 0.363 ms per 2^20 states)
#define SQMP_SINCOS_NOINLINE
#endif
#include "k1_scan.cuh"
#include "k1_self.cuh"

namespace sqmp {

#ifndef KF_THREADS_DEF
#define KF_THREADS_DEF 32
#endif
constexpr int KF_THREADS = KF_THREADS_DEF;
#ifndef KF_MIN_BLOCKS_DEF
#define KF_MIN_BLOCKS_DEF 18
#endif
constexpr int KF_MIN_BLOCKS = KF_MIN_BLOCKS_DEF;
constexpr float KF_MAGIC = 12582912.0f;          // 1.5 * 2^23: (x + MAGIC) has round-to-nearest-even(x) in its low mantissa bits
constexpr unsigned KF_MAGIC_BITS = 0x4B400000u;  // bit pattern of MAGIC

#ifdef SQMP_KF_STATS   // development builds only (tools/k1f_stats.py): where the work of the hierarchy goes
__device__ unsigned long long kf_stats[16];
#define KF_COUNT(slot, n) atomicAdd(&kf_stats[slot], (unsigned long long)(n))
#else
#define KF_COUNT(slot, n)
#endif

constexpr int KF_PQ_CAP = 32 + 8 * 32;   // a batch adds at most 8 x 32 entries to fewer than 32 left over
__host__ __device__ constexpr size_t kf_smem_base() {   // everything but the SELF frames
  return sizeof(WarpQueue) * (KF_THREADS / 32) + sizeof(double) * (SQMP_DOF + 12) * KF_THREADS + sizeof(float) * 12 * KF_THREADS +
         2 * sizeof(unsigned short) * KF_PQ_CAP * (KF_THREADS / 32);
}

__device__ __forceinline__ float kf_gap(int o, float f) {   // gap between coordinate f in [0,1] of the own cell and the unit cube at offset o
  return o > 0 ? float(o) - f : (o < 0 ? f - float(o) - 1.0f : 0.0f);
}

// Stage 1 for eight consecutive nodes (four pair records; `two` = false: only the first four exist): field bytes and
// thresholds.  F = frame rows 0-1 and the (possibly replaced) row 2, fp32; lanes with pred = false skip the gathers.
struct KfFrame {
  float r[9], p[3];
};
__device__ __forceinline__ void kf_stage1(const float4 *__restrict__ rec, bool two, const KfFrame &F, unsigned sx, unsigned spl,
                                          unsigned base_off, const uint8_t *__restrict__ lbp, bool pred, int (&lbv)[8], int (&thr)[8]) {
#pragma unroll
  for (int v = 0; v < 4; ++v) {   // pairs of nodes
    lbv[2 * v] = lbv[2 * v + 1] = 255;
    thr[2 * v] = thr[2 * v + 1] = -2;   // absent node / pad / disabled link: never open
    if (v < 2 || two) {
      const float4 A = rec[2 * v], B = rec[2 * v + 1];   // (x0 x1 y0 y1) (z0 z1 thr0 thr1)
      const float2 X = make_float2(A.x, A.y), Y = make_float2(A.z, A.w), Z = make_float2(B.x, B.y);
      const float2 cx = f32x2_fma(F.r[0], X, f32x2_fma(F.r[1], Y, f32x2_fma(F.r[2], Z, make_float2(F.p[0], F.p[0]))));
      const float2 cy = f32x2_fma(F.r[3], X, f32x2_fma(F.r[4], Y, f32x2_fma(F.r[5], Z, make_float2(F.p[1], F.p[1]))));
      const float2 cz = f32x2_fma(F.r[6], X, f32x2_fma(F.r[7], Y, f32x2_fma(F.r[8], Z, make_float2(F.p[2], F.p[2]))));
      const float2 tx = f32x2_add(cx, KF_MAGIC), ty = f32x2_add(cy, KF_MAGIC), tz = f32x2_add(cz, KF_MAGIC);
      const unsigned off0 = __float_as_uint(tz.x) * spl + __float_as_uint(ty.x) * sx + __float_as_uint(tx.x) + base_off;
      const unsigned off1 = __float_as_uint(tz.y) * spl + __float_as_uint(ty.y) * sx + __float_as_uint(tx.y) + base_off;
      if (pred) {
        lbv[2 * v] = int(lbp[off0]);
        lbv[2 * v + 1] = int(lbp[off1]);
      }
      thr[2 * v] = __float_as_int(B.z);
      thr[2 * v + 1] = __float_as_int(B.w);
    }
  }
}

// One-warp CTAs (21 per SM at 96 registers): a CTA lives as long as its slowest warp, so the smaller the CTA the shorter
// the tail — 128 -> 64 threads was worth 2.7 %, 64 -> 32 another 3.3 % (0.350 -> 0.338 ms back to back).  __launch_bounds__
// asks for 18 per SM: ptxas stays at 96 registers with 28 B of spills; asking for 21 (80 registers) spills 112 B for the same time.
// SELF keeps the group frames and the pair hierarchy live on top of the map state: 128 registers (512 threads per SM)
// avoid the spills that cost it 12 % at 96; the map-only kernel is indifferent between 512 and 640 threads per SM.
// SIG: compile-time chain signature (fk_device.cuh) — for the reference's robot every segment advance is straight-line
// code reached through one warp-uniform switch instead of the per-segment interpreter.
template <bool EDGES, bool SELF, typename SIG>
__global__ void __launch_bounds__(KF_THREADS, SELF ? (512 / KF_THREADS) : KF_MIN_BLOCKS)
k1_fast(const __grid_constant__ RobotDev rb, const __grid_constant__ GridDev gd, const __grid_constant__ SelfDev self,
        const __grid_constant__ FastDev fd, const __grid_constant__ KfTab tab, const __grid_constant__ PeerDev peer,
        const double *__restrict__ q_in, const double *__restrict__ q_end, int num_points, int64_t S,
        uint32_t *__restrict__ mask_out, uint32_t *__restrict__ first_invalid) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NW = KF_THREADS / 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  WarpQueue *wq = reinterpret_cast<WarpQueue *>(smem_raw) + wid;
  double *q_mine = reinterpret_cast<double *>(smem_raw + sizeof(WarpQueue) * NW) + threadIdx.x;  // [8][blockDim]
  // frames of the group the warp is working on, one column per lane: fp64 [12][32] (exact centres), fp32 [12][32] (stage 2)
  double *fr64 = reinterpret_cast<double *>(smem_raw + sizeof(WarpQueue) * NW + sizeof(double) * SQMP_DOF * KF_THREADS) + wid * (12 * 32);
  float *fr32 = reinterpret_cast<float *>(smem_raw + sizeof(WarpQueue) * NW + sizeof(double) * (SQMP_DOF + 12) * KF_THREADS) + wid * (12 * 32);
  // spheres stage 1 could not prove free, waiting for stage 2: (slot << 5 | owner lane), decided 32 at a time
  unsigned short *pq = reinterpret_cast<unsigned short *>(smem_raw + sizeof(WarpQueue) * NW + sizeof(double) * (SQMP_DOF + 12) * KF_THREADS +
                                                          sizeof(float) * 12 * KF_THREADS) + wid * KF_PQ_CAP;
  // clusters some state has to look into, waiting for stage 1 of their spheres: (first slot / 4) << 6 | eight slots << 5 | owner lane
  unsigned short *cq = pq + NW * KF_PQ_CAP;
  // SELF: the group frames [n_frames*12][blockDim]
  float *f_mine = reinterpret_cast<float *>(smem_raw + kf_smem_base()) + threadIdx.x;
  const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool active = s < S;
  const int64_t sl = active ? s : S - 1;  // idle lanes of the last warp shadow the last state (lockstep loops)

  // let the next grid of the stream be scheduled as soon as every CTA of this one has started: its first wave then
  // fills the SM slots this grid's last wave leaves idle (a no-op unless the launch opted in, see the launcher)
  asm volatile("griddepcontrol.launch_dependents;");

  // ---- the state (given, or interpolated from the edge end points), parked in shared memory (own column): the chain
  //      walk picks q[seg_qidx] with one LDS ----
  double q[SQMP_DOF];
  k1_load_state<EDGES>(q_in, q_end, num_points, sl, q);
  const bool sane = k1_park_state(rb, q, q_mine, KF_THREADS);
  bool hit = !sane;   // non-finite / absurd joint value: in collision (fail closed), no tests

  // ---- base reference point = position after the leading prismatic run; its cell bc ----
  double pref[3] = {rb.root_off[0], rb.root_off[1], rb.root_off[2]};
  for (int sg = 0; sg <= rb.ref_seg; ++sg) {
    const double qv = q_mine[rb.seg_qidx[sg] * KF_THREADS];
#pragma unroll
    for (int i = 0; i < 3; ++i) pref[i] += rb.seg_pre[sg][9 + i] + rb.seg_axis[sg][i] * qv;
  }
  Frame<double> f;
  frame_set_translation<double>(f, pref[0], pref[1], pref[2]);
  if (!sane) pref[0] = pref[1] = pref[2] = 0.0;
  int bc[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) bc[a] = rn_int(pref[a] - 0.5);
  // comfortable: the base lies inside the grid, so every node centre lies inside the halo-padded fields
  const bool comfy = sane && unsigned(bc[0]) < unsigned(gd.gx) && unsigned(bc[1]) < unsigned(gd.gy) && unsigned(bc[2]) < unsigned(gd.gz);
  const unsigned sx = gd.sx, splane = gd.splane, splane2 = gd.splane2;
  // linear offset of a cell from the raw bits of its three magic sums (all arithmetic mod 2^32; padded cells < 2^32);
  // a lane that is not comfortable gets a zeroed frame, so every one of its gathers reads the zero byte
  const unsigned base_off3 = (comfy ? gd.org_off + unsigned(bc[2]) * splane + unsigned(bc[1]) * sx + unsigned(bc[0]) : gd.zero_off) -
                             KF_MAGIC_BITS * (splane + sx + 1u);
  const unsigned base_off2 = (comfy ? gd.slab_off + unsigned(gd.halo) * sx + unsigned(gd.halo) + unsigned(bc[1]) * sx + unsigned(bc[0]) : gd.zero_off) -
                             KF_MAGIC_BITS * (splane2 + sx + 1u);
  const double bcd[3] = {double(bc[0]) + 0.5, double(bc[1]) + 0.5, double(bc[2]) + 0.5};
  const uint8_t *__restrict__ lbp = gd.lb2_base;

  int qcount = 0;  // warp-uniform
  for (int sg = rb.ref_seg + 1; sg < rb.n_seg; ++sg) {
    if constexpr (SIG::generic) {
      frame_advance_smem<double>(f, rb, sg, q_mine, KF_THREADS);
    } else {
      switch (sg) {
        case 0: frame_advance_sig<double, SIG, 0>(f, rb, q_mine, KF_THREADS); break;
        case 1: frame_advance_sig<double, SIG, 1>(f, rb, q_mine, KF_THREADS); break;
        case 2: frame_advance_sig<double, SIG, 2>(f, rb, q_mine, KF_THREADS); break;
        case 3: frame_advance_sig<double, SIG, 3>(f, rb, q_mine, KF_THREADS); break;
        case 4: frame_advance_sig<double, SIG, 4>(f, rb, q_mine, KF_THREADS); break;
        case 5: frame_advance_sig<double, SIG, 5>(f, rb, q_mine, KF_THREADS); break;
        case 6: frame_advance_sig<double, SIG, 6>(f, rb, q_mine, KF_THREADS); break;
        default: frame_advance_sig<double, SIG, 7>(f, rb, q_mine, KF_THREADS); break;
      }
    }
    const int g = sg + 1;
    const int bb = tab.grp_bound_end[g], be = tab.grp_bound_end[g + 1];
    if (bb == be && !(SELF && rb.grp_fidx[g] >= 0)) continue;
    // the group's frame rounded once to fp32, translation relative to the centre of the base's cell
    float R[9], pm[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = float(f.r[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) pm[k] = float(f.p[k] - bcd[k]);
    if (SELF) {
      const int fi = rb.grp_fidx[g];
      if (fi >= 0) {
        float *dst = f_mine + fi * 12 * KF_THREADS;
#pragma unroll
        for (int k = 0; k < 9; ++k) dst[k * KF_THREADS] = R[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) dst[(9 + k) * KF_THREADS] = pm[k];
      }
    }
    if (bb == be) continue;
    // stage-1 frame: zeroed for a lane outside the comfortable region; third row (0 0 1 | 0) for the 2-D group, whose
    // nodes carry their slab index as z
    const bool twod = g == tab.base2d_group;
    KfFrame F;
#pragma unroll
    for (int k = 0; k < 6; ++k) F.r[k] = comfy ? R[k] : 0.0f;
    F.r[6] = (comfy && !twod) ? R[6] : 0.0f;
    F.r[7] = (comfy && !twod) ? R[7] : 0.0f;
    F.r[8] = comfy ? (twod ? 1.0f : R[8]) : 0.0f;
    F.p[0] = comfy ? pm[0] : 0.0f;
    F.p[1] = comfy ? pm[1] : 0.0f;
    F.p[2] = (comfy && !twod) ? pm[2] : 0.0f;
    const unsigned spl = twod ? splane2 : splane;
    const unsigned base_off = twod ? base_off2 : base_off3;

    // ---- root bound of the group ----
    bool open;
    {
      const float4 rt = tab.root[g];
      const float cx = fmaf(F.r[0], rt.x, fmaf(F.r[1], rt.y, fmaf(F.r[2], rt.z, F.p[0])));
      const float cy = fmaf(F.r[3], rt.x, fmaf(F.r[4], rt.y, fmaf(F.r[5], rt.z, F.p[1])));
      const float cz = fmaf(F.r[6], rt.x, fmaf(F.r[7], rt.y, fmaf(F.r[8], rt.z, F.p[2])));
      const unsigned off = __float_as_uint(cz + KF_MAGIC) * spl + __float_as_uint(cy + KF_MAGIC) * sx + __float_as_uint(cx + KF_MAGIC) + base_off;
#ifdef SQMP_BOUNDS_CHECK   // debug builds only (tools/build_variant.sh): every gather must stay inside the fields
      if (off > gd.slab_off + unsigned(gd.n_slabs) * splane2) __trap();
#endif
      open = !hit && int(lbp[off]) <= __float_as_int(rt.w);
    }
    KF_COUNT(0, lane == 0);                                  // [0] group root tests (warp level)
    KF_COUNT(1, open);                                       // [1] open roots (lanes)
    if (!__any_sync(FULL, open)) continue;
    KF_COUNT(2, lane == 0);                                  // [2] groups entered (warp level)

    // park this group's frames where any lane can read them (stages 2-3 are decided by whichever lane is free)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      fr64[k * 32 + lane] = f.r[k];
      fr32[k * 32 + lane] = R[k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      fr64[(9 + k) * 32 + lane] = f.p[k];
      fr32[(9 + k) * 32 + lane] = pm[k];
    }
    __syncwarp();
    const unsigned comfy_mask = __ballot_sync(FULL, comfy);
    int pcount = 0;   // entries in pq (warp-uniform)

    // Stages 2-3 for 32 queued spheres at once, one per lane whatever state it belongs to: upper bound from the feature
    // field (certain hit), else the fp64 centre goes to the exact-scan queue.  Returns the owners proven in collision.
    auto resolve = [&](int base, int n) -> unsigned {
      const unsigned known = __ballot_sync(FULL, hit);
      bool amb = false, h2 = false;
      int i = 0, owner = lane;
      if (lane < n) {
        const unsigned e = pq[base + lane];
        owner = int(e & 31u);
        i = int(e >> 5);
      }
      const unsigned ob3 = __shfl_sync(FULL, base_off3, owner);   // the owner's field offset (its base cell)
      if (lane < n) {
        if (!((known >> owner) & 1u)) {
          if ((comfy_mask >> owner) & 1u) {   // stage 2: the true 3-D centre (fp32) against the feature field
            const float4 a = __ldg(reinterpret_cast<const float4 *>(fd.aux + i));   // x y z r2_lo
            const float *fo = fr32 + owner;
            const float cx = fmaf(fo[0 * 32], a.x, fmaf(fo[1 * 32], a.y, fmaf(fo[2 * 32], a.z, fo[9 * 32])));
            const float cy = fmaf(fo[3 * 32], a.x, fmaf(fo[4 * 32], a.y, fmaf(fo[5 * 32], a.z, fo[10 * 32])));
            const float cz = fmaf(fo[6 * 32], a.x, fmaf(fo[7 * 32], a.y, fmaf(fo[8 * 32], a.z, fo[11 * 32])));
            const float tx = cx + KF_MAGIC, ty = cy + KF_MAGIC, tz = cz + KF_MAGIC;
            const unsigned off = __float_as_uint(tz) * splane + __float_as_uint(ty) * sx + __float_as_uint(tx) + ob3;
#ifdef SQMP_BOUNDS_CHECK
            if (off >= gd.zero_off) __trap();
#endif
            const unsigned v = gd.ft_base[off];
            // position inside the own cell, in [0,1]: (c - n) is exact, c + 0.5 is the centre estimate
            const float fx = (cx - (tx - KF_MAGIC)) + 0.5f, fy = (cy - (ty - KF_MAGIC)) + 0.5f, fz = (cz - (tz - KF_MAGIC)) + 0.5f;
            const float gx_ = kf_gap(int(v & 31u) - 15, fx), gy_ = kf_gap(int((v >> 5) & 31u) - 15, fy),
                        gz_ = kf_gap(int((v >> 10) & 31u) - 15, fz);
            const float ub2 = fmaf(gx_, gx_, fmaf(gy_, gy_, gz_ * gz_));
            KF_COUNT(8, 1);                                  // [8] stage-2 tests (lanes)
            if (ub2 <= a.w) h2 = true;   // dist(fp64 centre, O*) <= rg: certain hit
            else amb = true;
          } else {
            amb = true;   // base outside the grid: reach test below, on the fp64 centre
          }
        }
      }
      double cx64 = 0.0, cy64 = 0.0, cz64 = 0.0, rg64 = 0.0;
      if (amb) {   // the fp64 centre exactly as K1 forms it, from the owner's fp64 frame
        const double2 *lp = reinterpret_cast<const double2 *>(fd.sph64 + i);
        const double2 l01 = __ldg(lp), l23 = __ldg(lp + 1);
        const double *fo = fr64 + owner;
        cx64 = fma(fo[0 * 32], l01.x, fma(fo[1 * 32], l01.y, fma(fo[2 * 32], l23.x, fo[9 * 32] - 0.5))) + 0.5;
        cy64 = fma(fo[3 * 32], l01.x, fma(fo[4 * 32], l01.y, fma(fo[5 * 32], l23.x, fo[10 * 32] - 0.5))) + 0.5;
        cz64 = fma(fo[6 * 32], l01.x, fma(fo[7 * 32], l01.y, fma(fo[8 * 32], l23.x, fo[11 * 32] - 0.5))) + 0.5;
        rg64 = l23.y;
        if (!((comfy_mask >> owner) & 1u)) {   // can the sphere still reach the grid?
          const int pad = __ldg(&fd.aux[i].ncell);
          const int jx = rn_int(cx64 - 0.5), jy = rn_int(cy64 - 0.5), jz = rn_int(cz64 - 0.5);
          amb = unsigned(jx + pad) < unsigned(gd.gx + 2 * pad) && unsigned(jy + pad) < unsigned(gd.gy + 2 * pad) &&
                unsigned(jz + pad) < unsigned(gd.gz + 2 * pad);
        }
      }
      unsigned hm = __reduce_or_sync(FULL, h2 ? (1u << owner) : 0u);
      const unsigned m = __ballot_sync(FULL, amb);
      if (m) {
        if (amb) {
          const int pos = qcount + __popc(m & ((1u << lane) - 1u));
          wq->x[pos] = cx64;
          wq->y[pos] = cy64;
          wq->z[pos] = cz64;
          wq->r[pos] = rg64;
          wq->info[pos] = (i << 8) | owner;
        }
        qcount += __popc(m);
        KF_COUNT(9, lane == 0 ? __popc(m) : 0);              // [9] exact-scan items
        if (qcount >= 32) {
          __syncwarp();
          KF_COUNT(10, lane == 0);                           // [10] queue drains
          hm |= drain_queue(gd, wq, qcount - 32, 32, known | hm);
          __syncwarp();
          qcount -= 32;
        }
      }
      return hm;
    };

    int ccount = 0;   // entries in cq (warp-uniform)
    // Stage 1 of the spheres of 32 queued clusters at once, one (cluster, state) per lane whatever state it belongs to:
    // the owner's frame comes from shared memory, the sphere records from global memory.  Undecided spheres go to pq.
    auto leaves_round = [&](int base, int n) {
      const unsigned known = __ballot_sync(FULL, hit);
      int owner = lane, i0 = 0;
      bool two = false, valid = false;
      if (lane < n) {
        const unsigned e = cq[base + lane];
        owner = int(e & 31u);
        two = (e >> 5) & 1u;
        i0 = int(e >> 6) << 2;
        valid = !((known >> owner) & 1u);
      }
      const unsigned boff = __shfl_sync(FULL, base_off, owner);   // the owner's field offset (its base cell; zero byte when not comfortable)
      unsigned pend = 0;
      if (valid) {
        const bool oc = (comfy_mask >> owner) & 1u;
        const float *fo = fr32 + owner;
        KfFrame G;   // the owner's stage-1 frame (see F above)
#pragma unroll
        for (int k = 0; k < 6; ++k) G.r[k] = oc ? fo[k * 32] : 0.0f;
        G.r[6] = (oc && !twod) ? fo[6 * 32] : 0.0f;
        G.r[7] = (oc && !twod) ? fo[7 * 32] : 0.0f;
        G.r[8] = oc ? (twod ? 1.0f : fo[8 * 32]) : 0.0f;
        G.p[0] = oc ? fo[9 * 32] : 0.0f;
        G.p[1] = oc ? fo[10 * 32] : 0.0f;
        G.p[2] = (oc && !twod) ? fo[11 * 32] : 0.0f;
        int lbv[8], thr[8];
        kf_stage1(fd.leafrec + 2 * (i0 >> 1), two, G, sx, spl, boff, lbp, true, lbv, thr);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (lbv[u] <= thr[u]) pend |= 1u << u;
      }
      if (!__any_sync(FULL, pend != 0)) return;
      const int cnt = __popc(pend);
      KF_COUNT(6, cnt);                                      // [6] spheres stage 1 cannot prove free (lanes)
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
      }
      int pos = pcount + incl - cnt;
      while (pend) {
        const int u = __ffs(pend) - 1;
        pend &= pend - 1;
        pq[pos++] = (unsigned short)(((i0 + u) << 5) | owner);
      }
      pcount += __shfl_sync(FULL, incl, 31);
      __syncwarp();
    };

    // a group with a single cluster: its root is that cluster's bound
    const bool single = be - bb == 8 && tab.bound_count[bb + 1] == 0;
    for (int b0 = bb; b0 < be; b0 += 8) {
      // ---- cluster bounds, eight at a time ----
      unsigned ob;
      {
        int lbv[8], thr[8];
        kf_stage1(tab.bound + 2 * (b0 >> 1), true, F, sx, spl, base_off, lbp, open && !hit && !single, lbv, thr);
        ob = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if ((single ? 0 : lbv[u]) <= thr[u]) ob |= 1u << u;   // (thr = -2: pad bound or disabled link)
        if (!open || hit) ob = 0;
      }
      unsigned todo = __reduce_or_sync(FULL, ob);
      KF_COUNT(3, lane == 0);                                // [3] bound batches (warp level)
      KF_COUNT(4, __popc(ob));                               // [4] open clusters (lanes)
      KF_COUNT(5, lane == 0 ? __popc(todo) : 0);             // [5] clusters entered (warp level)
      while (todo) {   // warp-uniform: queue (cluster, state) for every state that has to look into the cluster
        const int cb = __ffs(todo) - 1;
        todo &= todo - 1;
        const unsigned entry = (unsigned(tab.bound_first[b0 + cb]) >> 2) << 6 | (tab.bound_count[b0 + cb] > 4 ? 32u : 0u);
        const bool mine = (ob >> cb) & 1u;
        const unsigned m = __ballot_sync(FULL, mine);
        if (mine) cq[ccount + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(entry | unsigned(lane));
        ccount += __popc(m);
      }
      __syncwarp();
      // decide full rounds of 32; after the group's last bound batch whatever is left (the frames change with the next group)
      const bool last = b0 + 8 >= be;
      while (ccount >= 32 || (last && (ccount | pcount) != 0)) {
        const int n = min(ccount, 32);
        if (n > 0) {
          KF_COUNT(13, lane == 0);                           // [13] cluster rounds (warp level)
          leaves_round(ccount - n, n);
          ccount -= n;
        }
        const bool flush = last && ccount == 0;
        while (pcount >= 32 || (flush && pcount > 0)) {
          KF_COUNT(7, lane == 0);                            // [7] stage-2 rounds (warp level)
          const int np = min(pcount, 32);
          const unsigned hm = resolve(pcount - np, np);
          pcount -= np;
          if ((hm >> lane) & 1u) hit = true;
        }
      }
    }
    __syncwarp();
    // whole warp decided (CC:114-115: map before self, first hit wins) -> stop walking the chain
    if (!SELF && qcount == 0 && __all_sync(FULL, hit)) break;
  }
  if (qcount > 0) {
    __syncwarp();
    KF_COUNT(10, lane == 0);
    const unsigned hm = drain_queue(gd, wq, 0, qcount, __ballot_sync(FULL, hit));
    hit |= (hm >> lane) & 1u;
  }
  KF_COUNT(11, lane == 0);                                   // [11] warps
  KF_COUNT(12, hit);                                         // [12] colliding states

  if (SELF && !hit) hit = rb.self_const_collision || self_collision<false>(self, f_mine, KF_THREADS, nullptr);

  const unsigned word = __ballot_sync(FULL, hit && active);
  // Programmatic dependent launch: this grid may have started while the previous grid of the stream was still
  // draining its last wave (see the launcher).  It reads nothing a predecessor writes; its ONLY write is the mask, so
  // waiting here — until every prerequisite grid has completed and flushed — keeps back-to-back calls that reuse one
  // mask buffer equivalent to serial execution.  (A no-op when the grid was launched without the attribute.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (EDGES && first_invalid != nullptr && word != 0) {
    // isEdgeValid's early exit (BS:6866-6874) in the same pass: lane j looks at the j-th edge that overlaps this warp's 32
    // states and lowers first_invalid[e] (preset to 0xFFFFFFFF = -1 by the host) to the first colliding state it sees
    const int P1 = num_points + 1;
    const int64_t s0 = s - lane, e = s0 / P1 + lane;
    const int64_t b0 = max(e * P1, s0) - s0, b1 = min((e + 1) * P1, s0 + 32) - s0;
    if (b0 < b1) {
      const unsigned m = word & (b1 - b0 >= 32 ? ~0u : (((1u << int(b1 - b0)) - 1u) << int(b0)));
      if (m) atomicMin(first_invalid + e, unsigned(s0 + (__ffs(m) - 1) - e * P1));
    }
  }
  if (peer.n == 0) {
    if (lane == 0 && s < ((S + 31) & ~int64_t(31))) mask_out[s >> 5] = word;
    return;
  }
  // ---- fused all-gather: lane p stores the warp's word into rank p's gather buffer (one store instruction, n peers).
  //      Plain stores, no fence here: the completion of this grid orders them before the signal kernel that follows it ----
  if (lane < peer.n && s < ((S + 31) & ~int64_t(31))) peer.mask[lane][peer.word_offset + (s >> 5)] = word;
}

// after the validity kernel of a step (stream order, possibly through an event): publish the step in every rank's flags
__global__ void k_gather_signal(PeerDev peer) {
  __threadfence_system();
  if (threadIdx.x < peer.n) *reinterpret_cast<volatile uint32_t *>(peer.flag[threadIdx.x] + peer.rank) = peer.step;
}
cudaError_t launch_gather_signal(const PeerDev &peer, cudaStream_t st) {
  k_gather_signal<<<1, 32, 0, st>>>(peer);
  return cudaGetLastError();
}

// one warp: lane p waits until flags[p] has reached `step` (wrap-safe), bounded by ~2 s of SM clock
__global__ void k_gather_wait(const uint32_t *__restrict__ flags, int n, unsigned step, int32_t *__restrict__ status) {
  const int lane = threadIdx.x;
  bool ok = true;
  if (lane < n) {
    const long long t0 = clock64();
    const volatile uint32_t *f = flags + lane;
    while (int32_t(*f - step) < 0) {
      if (clock64() - t0 > 4000000000ll) { ok = false; break; }
      __nanosleep(200);
    }
  }
  __threadfence_system();   // acquire: the peers' mask words are visible to whatever follows on this stream
  const unsigned bad = __ballot_sync(0xffffffffu, !ok);
  if (lane == 0 && status) *status = bad ? 1 : 0;
}
cudaError_t launch_gather_wait(const uint32_t *d_flags, int n, unsigned step, int32_t *d_status, cudaStream_t st) {
  k_gather_wait<<<1, 32, 0, st>>>(d_flags, n, step, d_status);
  return cudaGetLastError();
}

#ifdef SQMP_KF_STATS
extern "C" int sqmp_debug_kf_stats(unsigned long long *out, int reset) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out, kf_stats, sizeof(kf_stats));
  if (e == cudaSuccess && reset) {
    unsigned long long z[16] = {0};
    e = cudaMemcpyToSymbol(kf_stats, z, sizeof(z));
  }
  return e == cudaSuccess ? 0 : -2;
}
#endif

cudaError_t launch_check_states_fast(const RobotDev &rb, const GridDev &gd, const SelfDev &self, const FastDev &fd,
                                     const KfTab &tab, const double *d_q, const double *d_q_end, int num_points, int64_t S,
                                     unsigned flags, uint32_t *d_mask, bool allow_specialized, bool overlap,
                                     cudaStream_t st, int *n_launch, const PeerDev *peers, int32_t *d_first_invalid) {
  const bool with_self = (flags & SQMP_CHECK_SELF) != 0;
  size_t smem = kf_smem_base();
  if (with_self) smem += size_t(12) * rb.n_frames * KF_THREADS * sizeof(float);
  const int64_t blocks = (S + KF_THREADS - 1) / KF_THREADS;
  if (blocks <= 0) return cudaSuccess;
  if (blocks > 0x7fffffff) return cudaErrorInvalidValue;
  cudaError_t err = cudaSuccess;
  PeerDev pd;
  std::memset(&pd, 0, sizeof(pd));
  if (peers) pd = *peers;
  auto go = [&](auto kern) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    // Shared memory / L1 split of the SM's unified array: the map-only kernel keeps 10 KB per one-warp CTA, so the default
    // carve-out (everything the 20 resident CTAs ask for) leaves ~50 KB of L1 for the field gathers (L1 hit 44 %).  At the
    // 164 KB step 16 CTAs stay resident and L1 gets ~85 KB: 3.3 % faster on config 2 and W512 (132 KB and below lose more
    // residency than the cache wins; the SELF variant needs its shared memory and keeps the default).
    if (!with_self) {
      int carve = 72;
#ifdef SQMP_KF_CARVEOUT_ENV
      if (const char *cv = getenv("SQMP_KF_CARVEOUT")) carve = atoi(cv);
#endif
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(blocks));
    cfg.blockDim = dim3(KF_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // may overlap the tail of the preceding grid
    attr[0].val.programmaticStreamSerializationAllowed = overlap ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    err = cudaLaunchKernelEx(&cfg, kern, rb, gd, self, fd, tab, pd, d_q, d_q_end, d_q_end ? num_points : 0, S, d_mask,
                             reinterpret_cast<uint32_t *>(d_first_invalid));
  };
  if (allow_specialized && chain_matches<SigRobotino>(rb)) {   // structure of the chain known at compile time (the reference's robot)
    if (d_q_end) {
      if (with_self) go(k1_fast<true, true, SigRobotino>);
      else go(k1_fast<true, false, SigRobotino>);
    } else {
      if (with_self) go(k1_fast<false, true, SigRobotino>);
      else go(k1_fast<false, false, SigRobotino>);
    }
  } else if (d_q_end) {
    if (with_self) go(k1_fast<true, true, SigGeneric>);
    else go(k1_fast<true, false, SigGeneric>);
  } else {
    if (with_self) go(k1_fast<false, true, SigGeneric>);
    else go(k1_fast<false, false, SigGeneric>);
  }
  if (n_launch) *n_launch += 1;
  return err != cudaSuccess ? err : cudaGetLastError();
}

}  // namespace sqmp
