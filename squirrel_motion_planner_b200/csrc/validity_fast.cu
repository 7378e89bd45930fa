// validity_fast.cu — K1F: the state-validity kernel the library uses whenever its preconditions hold (sqmp_create /
// sqmp_set_voxel_grid decide; otherwise the generic K1 of validity.cu runs).  Same predicate, same verdicts.
//
// Reference path replaced (file:line in /root/reference): CollisionChecker::isInCollision CC:104-121, updateTransforms
// CC:519-539, checkMapCollision CC:554-592, checkSelfCollision CC:541-552; BiRRTstarPlanner::interpolateConfigurations
// BS:4443-4490, isEdgeValid BS:6864-6895.
//
// Why a second kernel.  The generic K1 spends ~45 % of its instructions proving "this sphere is nowhere near an
// obstacle" in fp64 with per-sphere bounds logic.  Here that proof runs in fp32 and is still exact:
//   * the chain is walked in fp64 (identical frames to K1); per rigid group the frame is rounded ONCE to fp32 in
//     coordinates relative to the base's own cell, so |fp32 centre - fp64 centre| <= err (host-computed bound, ~5e-5
//     cells; capi.cu: refresh_grid_robot);
//   * stage 1 (every sphere): two spheres per packed fp32x2 instruction (9 FFMA2 + 3 FADD2 per pair), 2 IMAD + 1 add per
//     sphere for the cell offset, one byte gather from the halo-padded lower-bound field — no bounds test, because the
//     field covers every cell a sphere can reach while the base is inside the grid.
//     "lb2 > thr" with thr = floor((rg + err)^2) proves the fp64 centre free as well;
//   * stage 2 (the ~11 % that stage 1 cannot prove free): the feature-field upper bound in fp32 against
//     (rg - err)^2 proves a hit for the fp64 centre as well;
//   * stage 3 (neither bound decides, ~0.4 spheres per state after the early exits): the fp64 centre is formed from the
//     fp64 frame exactly as K1 forms it and queued for the same exact bit-row scan (k1_scan.cuh).
// Grids are launched with programmatic stream serialization (see the launcher): the next grid of the stream may start
// while this one drains, and writes its mask only after this one has completed.
// States whose base lies outside the grid (or is not finite) never use the fp32 path: their spheres go through the
// reach test and the exact scan only.
#include <cstring>

#include "k1_scan.cuh"
#include "k1_self.cuh"

namespace sqmp {

#ifndef KF_THREADS_DEF
#define KF_THREADS_DEF 64
#endif
constexpr int KF_THREADS = KF_THREADS_DEF;
#ifndef KF_MIN_BLOCKS_DEF
#define KF_MIN_BLOCKS_DEF 10
#endif
constexpr int KF_MIN_BLOCKS = KF_MIN_BLOCKS_DEF;
constexpr float KF_MAGIC = 12582912.0f;          // 1.5 * 2^23: (x + MAGIC) has round-to-nearest-even(x) in its low mantissa bits
constexpr unsigned KF_MAGIC_BITS = 0x4B400000u;  // bit pattern of MAGIC

__device__ __forceinline__ float kf_gap(int o, float f) {   // gap between coordinate f in [0,1] of the own cell and the unit cube at offset o
  return o > 0 ? float(o) - f : (o < 0 ? f - float(o) - 1.0f : 0.0f);
}

// 64-thread CTAs (10 per SM at 96 registers): the shorter tail of small CTAs is worth 2.7 % over 128-thread ones (measured).
// SELF keeps the group frames and the pair hierarchy live on top of the map state: 128 registers (512 threads per SM)
// avoid the spills that cost it 12 % at 96; the map-only kernel is indifferent between 512 and 640 threads per SM.
// SIG: compile-time chain signature (fk_device.cuh) — for the reference's robot every segment advance is straight-line
// code reached through one warp-uniform switch instead of the per-segment interpreter.
template <bool EDGES, bool SELF, typename SIG>
__global__ void __launch_bounds__(KF_THREADS, SELF ? (512 / KF_THREADS) : KF_MIN_BLOCKS)
k1_fast(const __grid_constant__ RobotDev rb, const __grid_constant__ GridDev gd, const __grid_constant__ SelfDev self,
        const __grid_constant__ FastDev fd, const __grid_constant__ PeerDev peer, const double *__restrict__ q_in,
        const double *__restrict__ q_end, int num_points, int64_t S, uint32_t *__restrict__ mask_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpQueue *wq = reinterpret_cast<WarpQueue *>(smem_raw) + (threadIdx.x >> 5);
  double *q_mine = reinterpret_cast<double *>(smem_raw + sizeof(WarpQueue) * (KF_THREADS / 32)) + threadIdx.x;  // [8][blockDim]
  float4 *slots4 = reinterpret_cast<float4 *>(smem_raw + sizeof(WarpQueue) * (KF_THREADS / 32) + sizeof(double) * SQMP_DOF * KF_THREADS);
  // slots4: [n_slots / 2] FastPair (2 x float4 each), then [n_slots] FastAux (8 bytes each)
  const FastAux *aux_s = reinterpret_cast<const FastAux *>(slots4 + fd.n_slots);
  // SELF: copies of the block and A-sphere tables of the pair hierarchy, then the group frames [n_frames*12][blockDim]
  SelfBlock *blocks_s = reinterpret_cast<SelfBlock *>(slots4 + fd.n_slots + (fd.n_slots + 1) / 2);
  SelfA *as_s = reinterpret_cast<SelfA *>(blocks_s + (SELF ? self.n_blocks : 0));
  float *f_mine = reinterpret_cast<float *>(as_s + (SELF ? self.n_as : 0)) + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool active = s < S;
  const int64_t sl = active ? s : S - 1;  // idle lanes of the last warp shadow the last state (lockstep loops)

  // let the next grid of the stream be scheduled as soon as every CTA of this one has started: its first wave then
  // fills the SM slots this grid's last wave leaves idle (one CTA lifetime, ~36 us of a 2^20-state launch)
  asm volatile("griddepcontrol.launch_dependents;");

  // ---- slot tables -> shared memory (16 + 8 bytes per slot) ----
  {
    const float4 *src = reinterpret_cast<const float4 *>(fd.pairs);
    for (int i = threadIdx.x; i < fd.n_slots; i += KF_THREADS) slots4[i] = __ldg(src + i);
    const float2 *src2 = reinterpret_cast<const float2 *>(fd.aux);
    float2 *dst2 = reinterpret_cast<float2 *>(slots4 + fd.n_slots);
    for (int i = threadIdx.x; i < fd.n_slots; i += KF_THREADS) dst2[i] = __ldg(src2 + i);
    if (SELF) {
      const float4 *sb = reinterpret_cast<const float4 *>(self.blocks), *sa = reinterpret_cast<const float4 *>(self.as);
      float4 *db = reinterpret_cast<float4 *>(blocks_s), *da = reinterpret_cast<float4 *>(as_s);
      for (int i = threadIdx.x; i < 3 * self.n_blocks; i += KF_THREADS) db[i] = __ldg(sb + i);
      for (int i = threadIdx.x; i < 2 * self.n_as; i += KF_THREADS) da[i] = __ldg(sa + i);
    }
  }

  // ---- the state (given, or interpolated from the edge end points), parked in shared memory (own column): the chain
  //      walk picks q[seg_qidx] with one LDS ----
  double q[SQMP_DOF];
  k1_load_state<EDGES>(q_in, q_end, num_points, sl, q);
  const bool sane = k1_park_state(rb, q, q_mine, KF_THREADS);   // false: non-finite / absurd joint value, no map test

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
  // comfortable: the base lies inside the grid, so every sphere centre lies inside the halo-padded fields
  const bool comfy = sane && unsigned(bc[0]) < unsigned(gd.gx) && unsigned(bc[1]) < unsigned(gd.gy) && unsigned(bc[2]) < unsigned(gd.gz);
  const unsigned sx = gd.sx, splane = gd.splane;
  // linear offset of a cell from the raw bits of its three magic sums (all arithmetic mod 2^32; padded cells < 2^32)
  const unsigned base_off = comfy ? gd.org_off + unsigned(bc[2]) * splane + unsigned(bc[1]) * sx + unsigned(bc[0]) -
                                        KF_MAGIC_BITS * (splane + sx + 1u)
                                  : gd.zero_off - KF_MAGIC_BITS * (splane + sx + 1u);   // frame zeroed below: every gather reads the zero byte
  const double bcd[3] = {double(bc[0]) + 0.5, double(bc[1]) + 0.5, double(bc[2]) + 0.5};

  __syncthreads();   // slot table staged

  bool hit = false;
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
    const int b = rb.seg_sph_end[sg + 1], e = rb.seg_sph_end[sg + 2];
    if (b == e) continue;
    // the group's frame rounded once to fp32, translation relative to the centre of the base's cell
    float R[9], pm[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = float(f.r[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) pm[k] = float(f.p[k] - bcd[k]);
    if (SELF) {
      const int fi = rb.grp_fidx[sg + 1];
      if (fi >= 0) {
        float *dst = f_mine + fi * 12 * KF_THREADS;
#pragma unroll
        for (int k = 0; k < 9; ++k) dst[k * KF_THREADS] = R[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) dst[(9 + k) * KF_THREADS] = pm[k];
      }
    }
    if (!comfy) {   // every centre -> offset 0 -> the zero byte: lb2 = 0 <= thr keeps every enabled sphere pending
#pragma unroll
      for (int k = 0; k < 9; ++k) R[k] = 0.0f;
      pm[0] = pm[1] = pm[2] = 0.0f;
    }
    const uint8_t *__restrict__ lbp = gd.lb2_base;
    for (int i0 = b; i0 < e; i0 += 8) {
      const bool two = i0 + 4 < e;   // warp-uniform: slots come in multiples of 4
      int lbv[8], thr[8];
#pragma unroll
      for (int v = 0; v < 4; ++v) {   // pairs of slots
        lbv[2 * v] = lbv[2 * v + 1] = 255;
        thr[2 * v] = thr[2 * v + 1] = -2;   // absent slot / pad / disabled link: never pending
        if (v < 2 || two) {
          const float4 A = slots4[i0 + 2 * v], B = slots4[i0 + 2 * v + 1];   // (x0 x1 y0 y1) (z0 z1 thr0 thr1)
          const float2 X = make_float2(A.x, A.y), Y = make_float2(A.z, A.w), Z = make_float2(B.x, B.y);
          const float2 cx = f32x2_fma(R[0], X, f32x2_fma(R[1], Y, f32x2_fma(R[2], Z, make_float2(pm[0], pm[0]))));
          const float2 cy = f32x2_fma(R[3], X, f32x2_fma(R[4], Y, f32x2_fma(R[5], Z, make_float2(pm[1], pm[1]))));
          const float2 cz = f32x2_fma(R[6], X, f32x2_fma(R[7], Y, f32x2_fma(R[8], Z, make_float2(pm[2], pm[2]))));
          const float2 tx = f32x2_add(cx, KF_MAGIC), ty = f32x2_add(cy, KF_MAGIC), tz = f32x2_add(cz, KF_MAGIC);
          const unsigned off0 = __float_as_uint(tz.x) * splane + __float_as_uint(ty.x) * sx + __float_as_uint(tx.x) + base_off;
          const unsigned off1 = __float_as_uint(tz.y) * splane + __float_as_uint(ty.y) * sx + __float_as_uint(tx.y) + base_off;
#ifdef SQMP_BOUNDS_CHECK   // debug builds only (tools/build_variant.sh): every gather must stay inside the padded field
          if (off0 > gd.zero_off || off1 > gd.zero_off) __trap();
#endif
          lbv[2 * v] = int(lbp[off0]);
          lbv[2 * v + 1] = int(lbp[off1]);
          thr[2 * v] = __float_as_int(B.z);
          thr[2 * v + 1] = __float_as_int(B.w);
        }
      }
      unsigned pend = 0;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (lbv[u] <= thr[u]) pend |= 1u << u;
      if (hit || !sane) pend = 0;
      while (__any_sync(FULL, pend != 0)) {
        bool amb = false;
        int i = i0;
        double cx64 = 0.0, cy64 = 0.0, cz64 = 0.0, rg64 = 0.0;
        auto centre64 = [&]() {   // the fp64 centre exactly as K1 forms it
          const double2 *lp = reinterpret_cast<const double2 *>(fd.sph64 + i);
          const double2 l01 = __ldg(lp), l23 = __ldg(lp + 1);
          const struct { double x, y, z, w; } l = {l01.x, l01.y, l23.x, l23.y};
          cx64 = fma(f.r[0], l.x, fma(f.r[1], l.y, fma(f.r[2], l.z, f.p[0] - 0.5))) + 0.5;
          cy64 = fma(f.r[3], l.x, fma(f.r[4], l.y, fma(f.r[5], l.z, f.p[1] - 0.5))) + 0.5;
          cz64 = fma(f.r[6], l.x, fma(f.r[7], l.y, fma(f.r[8], l.z, f.p[2] - 0.5))) + 0.5;
          rg64 = l.w;
        };
        if (pend != 0) {
          const int u = __ffs(pend) - 1;
          pend &= pend - 1;
          i = i0 + u;
          if (comfy) {
            const float *sp = reinterpret_cast<const float *>(slots4) + 8 * (i >> 1) + (i & 1);
            const float3 a = make_float3(sp[0], sp[2], sp[4]);
            const float r2lo = aux_s[i].r2_lo;
            const float cx = fmaf(R[0], a.x, fmaf(R[1], a.y, fmaf(R[2], a.z, pm[0])));
            const float cy = fmaf(R[3], a.x, fmaf(R[4], a.y, fmaf(R[5], a.z, pm[1])));
            const float cz = fmaf(R[6], a.x, fmaf(R[7], a.y, fmaf(R[8], a.z, pm[2])));
            const float tx = cx + KF_MAGIC, ty = cy + KF_MAGIC, tz = cz + KF_MAGIC;
            const unsigned off = __float_as_uint(tz) * splane + __float_as_uint(ty) * sx + __float_as_uint(tx) + base_off;
#ifdef SQMP_BOUNDS_CHECK
            if (off >= gd.zero_off) __trap();
#endif
            const unsigned v = gd.ft_base[off];
            // position inside the own cell, in [0,1]: (c - n) is exact, c + 0.5 is the centre estimate
            const float fx = (cx - (tx - KF_MAGIC)) + 0.5f, fy = (cy - (ty - KF_MAGIC)) + 0.5f, fz = (cz - (tz - KF_MAGIC)) + 0.5f;
            const float gx_ = kf_gap(int(v & 31u) - 15, fx), gy_ = kf_gap(int((v >> 5) & 31u) - 15, fy),
                        gz_ = kf_gap(int((v >> 10) & 31u) - 15, fz);
            const float ub2 = fmaf(gx_, gx_, fmaf(gy_, gy_, gz_ * gz_));
            if (ub2 <= r2lo) { hit = true; pend = 0; }   // dist(fp64 centre, O*) <= rg: certain hit
            else amb = true;
          } else {   // base outside the grid: can the sphere still reach it?
            centre64();
            const int pad = aux_s[i].ncell;
            const int jx = rn_int(cx64 - 0.5), jy = rn_int(cy64 - 0.5), jz = rn_int(cz64 - 0.5);
            amb = unsigned(jx + pad) < unsigned(gd.gx + 2 * pad) && unsigned(jy + pad) < unsigned(gd.gy + 2 * pad) &&
                  unsigned(jz + pad) < unsigned(gd.gz + 2 * pad);
          }
        }
        const unsigned m = __ballot_sync(FULL, amb);
        if (m) {
          if (amb) {
            if (comfy) centre64();
            const int pos = qcount + __popc(m & ((1u << lane) - 1u));
            wq->x[pos] = cx64;
            wq->y[pos] = cy64;
            wq->z[pos] = cz64;
            wq->r[pos] = rg64;
            wq->info[pos] = (i << 8) | lane;
          }
          qcount += __popc(m);
          if (qcount >= 32) {
            __syncwarp();
            const unsigned hm = drain_queue(gd, wq, qcount - 32, 32, __ballot_sync(FULL, hit));
            __syncwarp();
            qcount -= 32;
            if ((hm >> lane) & 1u) { hit = true; pend = 0; }
          }
        }
      }
    }
    // whole warp decided (CC:114-115: map before self, first hit wins) -> stop walking the chain
    if (!SELF && qcount == 0 && __all_sync(FULL, hit)) break;
  }
  if (qcount > 0) {
    __syncwarp();
    const unsigned hm = drain_queue(gd, wq, 0, qcount, __ballot_sync(FULL, hit));
    hit |= (hm >> lane) & 1u;
  }

  if (SELF && !hit)
    hit = rb.self_const_collision || self_collision(blocks_s, self.n_blocks, as_s, self.hpairs, f_mine, KF_THREADS);

  const unsigned word = __ballot_sync(FULL, hit && active);
  // Programmatic dependent launch: this grid may have started while the previous grid of the stream was still
  // draining its last wave (see the launcher).  It reads nothing a predecessor writes; its ONLY write is the mask, so
  // waiting here — until every prerequisite grid has completed and flushed — keeps back-to-back calls that reuse one
  // mask buffer equivalent to serial execution.  (A no-op when the grid was launched without the attribute.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
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

cudaError_t launch_check_states_fast(const RobotDev &rb, const GridDev &gd, const SelfDev &self, const FastDev &fd,
                                     const double *d_q, const double *d_q_end, int num_points, int64_t S,
                                     unsigned flags, uint32_t *d_mask, bool allow_specialized, bool overlap,
                                     cudaStream_t st, int *n_launch, const PeerDev *peers) {
  const bool with_self = (flags & SQMP_CHECK_SELF) != 0;
  size_t smem = sizeof(WarpQueue) * (KF_THREADS / 32) + sizeof(double) * SQMP_DOF * KF_THREADS +
                16 * size_t(fd.n_slots + (fd.n_slots + 1) / 2);   // FastPair: 16 B per slot; FastAux: 8 B per slot
  if (with_self)
    smem += size_t(12) * rb.n_frames * KF_THREADS * sizeof(float) + sizeof(SelfBlock) * self.n_blocks + sizeof(SelfA) * self.n_as;
  const int64_t blocks = (S + KF_THREADS - 1) / KF_THREADS;
  if (blocks <= 0) return cudaSuccess;
  if (blocks > 0x7fffffff) return cudaErrorInvalidValue;
  cudaError_t err = cudaSuccess;
  PeerDev pd;
  std::memset(&pd, 0, sizeof(pd));
  if (peers) pd = *peers;
  auto go = [&](auto kern) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
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
    err = cudaLaunchKernelEx(&cfg, kern, rb, gd, self, fd, pd, d_q, d_q_end, d_q_end ? num_points : 0, S, d_mask);
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
