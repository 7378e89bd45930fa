// fk.cu — K2: batched end-effector forward kinematics, 64 B in / 56 B out per state (the HBM-bound kernel).
//
// Reference path replaced: KDLRobotModel::compute_FK (kdl_kuka_model.cpp:278-305), which copies the KDL chain by
// value and constructs a ChainFkSolverPos_recursive for every single call, as invoked 21x per candidate edge by
// BiRRTstarPlanner::interpolateConfigurations (birrt_star.cpp:4486-4487).
//
// Data movement: a CTA owns a tile of FK_TILE consecutive states.  The 64-byte input rows and 56-byte output
// rows are contiguous in global memory, so the tile is moved with 1-D bulk async copies (TMA, cp.async.bulk ->
// SASS UBLKCP) between global and shared memory: one elected thread issues the load and waits on an mbarrier,
// every thread then reads its own row from shared memory, computes the chain in registers, writes its 7 doubles
// back to shared memory, and one elected thread issues the bulk store.  Global traffic is therefore exactly the
// algorithmic 120 B/state in full 128-byte lines, with no per-thread strided LDG/STG.
#include "fk_device.cuh"

namespace sqmp {

constexpr int FK_TILE = 128;  // states per CTA iteration == threads per CTA
// row pitches in shared memory (doubles).  Input rows keep the global pitch (8) because the bulk copy is a flat
// byte copy; each joint value is read from the tile exactly once, when its chain segment needs it (the 8-way bank
// conflict of that strided LDS.64 costs ~60 cycles per warp against ~900 instructions of chain arithmetic).
// Outputs use pitch 7 (odd) and are conflict-free as they are.
constexpr int FK_IN_PITCH = 8, FK_OUT_PITCH = 7;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// The chain for one state, its joint vector read straight from the shared-memory tile row (no register copy):
// T = double everywhere, or T = float for the rotational part with the prismatic base added in double (see DESIGN.md
// for the error budget against the 1e-5 m / 1e-5 rad tolerance).
template <typename T, typename SIG, bool SMALL = false>
__device__ __forceinline__ void fk_one(const RobotDev &rb, const double *q_row, double *out7) {
  Frame<T> f;
  frame_set_translation<T>(f, T(0), T(0), T(0));
  double off[3] = {0.0, 0.0, 0.0};
  if constexpr (SIG::generic) {
    int sg0 = 0;
    if (sizeof(T) == 4) {
      for (int sg = 0; sg <= rb.ref_seg; ++sg) {
        const double qv = q_row[rb.seg_qidx[sg]];
#pragma unroll
        for (int i = 0; i < 3; ++i) off[i] += rb.seg_pre[sg][9 + i] + rb.seg_axis[sg][i] * qv;
      }
      sg0 = rb.ref_seg + 1;
    }
    for (int sg = sg0; sg < rb.n_seg; ++sg) frame_advance_smem<T>(f, rb, sg, q_row, 1);
    if (sizeof(T) == 4) frame_mul_const<T>(f, rb.ee_tip_f, rb.ee_tip_kind);
    else frame_mul_const<T>(f, rb.ee_tip, rb.ee_tip_kind);
  } else {
    // leading prismatic x, y (identity pre-frames, signature): a plain fp64 offset in both precisions
    off[0] = q_row[0];
    off[1] = q_row[1];
    chain_walk_sig<T, SIG, 2, SIG::n, SMALL>(f, rb, q_row, 1);
    if (sizeof(T) == 4) frame_mul_const<T>(f, rb.ee_tip_f, SIG::ee_kind);
    else frame_mul_const<T>(f, rb.ee_tip, SIG::ee_kind);
  }
  double R[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = double(f.r[i]);
  out7[0] = double(f.p[0]) + off[0];
  out7[1] = double(f.p[1]) + off[1];
  out7[2] = double(f.p[2]) + off[2];
  get_quaternion(R, out7[3], out7[4], out7[5], out7[6]);
}

// Warp-private tiles: every warp of the CTA streams its own 32-state tiles (2 KB in, 1.75 KB out) through its own
// double-buffered shared-memory slice with its own pair of mbarriers — no __syncthreads anywhere, so a warp never waits
// for the slowest warp of its CTA (the CTA-wide barriers of the previous version held ~14 % of the stall samples).
#ifndef FK_ROWS_PER_LANE_DEF
#define FK_ROWS_PER_LANE_DEF 2
#endif
constexpr int FK_RPL = FK_ROWS_PER_LANE_DEF;   // states per lane: two independent chains per thread keep the fp64 pipe fed
constexpr int FK_WARP_TILE = 32 * FK_RPL;
#ifndef FK_MIN_BLOCKS_DEF
#define FK_MIN_BLOCKS_DEF 4
#endif
template <typename T, typename SIG>
__global__ void __launch_bounds__(FK_TILE, FK_MIN_BLOCKS_DEF)
k2_fk(const __grid_constant__ RobotDev rb, const double *__restrict__ q_in, int64_t S, double *__restrict__ out) {
  constexpr int WARPS = FK_TILE / 32;
  __shared__ __align__(128) double s_in[WARPS][2][FK_WARP_TILE * FK_IN_PITCH];
  __shared__ __align__(128) double s_out[WARPS][FK_WARP_TILE * FK_OUT_PITCH];
  __shared__ __align__(8) uint64_t bar[WARPS][2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t n_tiles = (S + FK_WARP_TILE - 1) / FK_WARP_TILE;
  const int64_t stride = int64_t(gridDim.x) * WARPS;
  if (lane == 0) {
    mbar_init(&bar[w][0], 1);
    mbar_init(&bar[w][1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  auto tile_rows = [&](int64_t tile) { return int(imin64(FK_WARP_TILE, S - tile * FK_WARP_TILE)); };
  int64_t tile = int64_t(blockIdx.x) * WARPS + w;
  if (lane == 0 && tile < n_tiles) {
    const uint32_t bytes = uint32_t(tile_rows(tile)) * SQMP_DOF * sizeof(double);
    mbar_expect_tx(&bar[w][0], bytes);
    bulk_g2s(s_in[w][0], q_in + tile * FK_WARP_TILE * SQMP_DOF, bytes, &bar[w][0]);
  }
  uint32_t phase[2] = {0, 0};
  int buf = 0;
  for (; tile < n_tiles; tile += stride, buf ^= 1) {
    // prefetch this warp's next tile into its other buffer (consumed one iteration ago, before the __syncwarp below)
    const int64_t next = tile + stride;
    if (lane == 0 && next < n_tiles) {
      const uint32_t bytes = uint32_t(tile_rows(next)) * SQMP_DOF * sizeof(double);
      mbar_expect_tx(&bar[w][buf ^ 1], bytes);
      bulk_g2s(s_in[w][buf ^ 1], q_in + next * FK_WARP_TILE * SQMP_DOF, bytes, &bar[w][buf ^ 1]);
    }
    mbar_wait(&bar[w][buf], phase[buf]);
    phase[buf] ^= 1;

    const int rows = tile_rows(tile);
    if (rows == FK_WARP_TILE) {
      // full tile: every lane owns FK_RPL rows (lane, lane + 32, ...).  When all their joint values are small enough for
      // the branch-free sincos the chains are straight-line code and interleave; otherwise row by row.
      const double *rows_in = s_in[w][buf];
      bool small = true;
      if constexpr (!SIG::generic && sizeof(T) == 8) {
#pragma unroll
        for (int r = 0; r < FK_RPL; ++r)
#pragma unroll
          for (int j = 2; j < SQMP_DOF; ++j) small = small && fabs(rows_in[(lane + 32 * r) * FK_IN_PITCH + j]) < 1.0e5;
      } else {
        small = false;
      }
      if (small) {
        double o[FK_RPL][7];
#pragma unroll
        for (int r = 0; r < FK_RPL; ++r) fk_one<T, SIG, true>(rb, rows_in + (lane + 32 * r) * FK_IN_PITCH, o[r]);
#pragma unroll
        for (int r = 0; r < FK_RPL; ++r) {
          double *orow = s_out[w] + (lane + 32 * r) * FK_OUT_PITCH;
#pragma unroll
          for (int j = 0; j < 7; ++j) orow[j] = o[r][j];
        }
      } else {
#pragma unroll 1
        for (int r = 0; r < FK_RPL; ++r) {
          double o[7];
          fk_one<T, SIG>(rb, rows_in + (lane + 32 * r) * FK_IN_PITCH, o);
          double *orow = s_out[w] + (lane + 32 * r) * FK_OUT_PITCH;
#pragma unroll
          for (int j = 0; j < 7; ++j) orow[j] = o[j];
        }
      }
    } else {
      // ragged last tile: its byte count need not be a multiple of 16, which bulk copies require
      for (int row = lane; row < rows; row += 32) {
        double o[7];
        fk_one<T, SIG>(rb, s_in[w][buf] + row * FK_IN_PITCH, o);
        double *orow = out + (tile * FK_WARP_TILE + row) * 7;
#pragma unroll
        for (int j = 0; j < 7; ++j) orow[j] = o[j];
      }
    }
    fence_async_smem();  // make the generic-proxy writes to s_out visible to the async (bulk copy) proxy
    __syncwarp();
    if (lane == 0 && rows == FK_WARP_TILE) {
      bulk_s2g(out + tile * FK_WARP_TILE * 7, s_out[w], uint32_t(FK_WARP_TILE) * 7 * sizeof(double));
      bulk_commit();
      bulk_wait_read0();  // s_out[w] may be rewritten once the copy engine has read it
    }
    __syncwarp();
  }
  if (lane == 0) bulk_wait0();
}

cudaError_t launch_fk(const RobotDev &rb, const double *d_q, int64_t S, double *d_out, int fk_precision,
                      bool allow_specialized, cudaStream_t st) {
  if (S <= 0) return cudaSuccess;
  static int sms = 0;   // one device kind per process (B200)
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int64_t n_tiles = (S + FK_TILE * FK_RPL - 1) / (FK_TILE * FK_RPL);   // CTAs needed if every warp took one tile
  // persistent CTAs: a multiple of the SM count, enough resident warps to cover the fp64 latency
  const unsigned blocks = unsigned(imin64(n_tiles, int64_t(sms) * FK_MIN_BLOCKS_DEF));
  const bool robotino = allow_specialized && chain_matches<SigRobotino>(rb);   // structure known at compile time: unrolled chain
  if (fk_precision == 32) {
    if (robotino) k2_fk<float, SigRobotino><<<blocks, FK_TILE, 0, st>>>(rb, d_q, S, d_out);
    else k2_fk<float, SigGeneric><<<blocks, FK_TILE, 0, st>>>(rb, d_q, S, d_out);
  } else {
    if (robotino) k2_fk<double, SigRobotino><<<blocks, FK_TILE, 0, st>>>(rb, d_q, S, d_out);
    else k2_fk<double, SigGeneric><<<blocks, FK_TILE, 0, st>>>(rb, d_q, S, d_out);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// K5: batched configuration sampling — sampleJointConfig_JntArray (birrt_star.cpp:3832-3878) over getRandomConf
// (control_laws.cpp:1120-1188): uniform in the joint / environment limits, redrawn until the end-effector lies above
// the platform (0.0 <= ee_z).  One thread per sample; the draw is a pure function of (seed, sample index, try, joint)
// through the splitmix64 stream (k-th output = mix(seed + (k + 1) * GAMMA)), so batches are reproducible and
// identical to the CPU oracle.  a + (b - a) * u is evaluated with explicit round-to-nearest mul/add (no FMA).
// ---------------------------------------------------------------------------------------------------------------
struct SampleBox {
  double lo[SQMP_DOF], hi[SQMP_DOF];
};
__device__ __forceinline__ uint64_t splitmix_at(uint64_t seed, uint64_t k) {
  uint64_t z = seed + (k + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <typename SIG>
__global__ void __launch_bounds__(128)
k5_sample(const __grid_constant__ RobotDev rb, const __grid_constant__ SampleBox box, uint64_t seed, int64_t first,
          int64_t n, int max_tries, double *__restrict__ q_out, int32_t *__restrict__ tries_out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double q[SQMP_DOF], o[7];
  int used = -max_tries;
  for (int t = 0; t < max_tries; ++t) {
    const uint64_t k0 = (uint64_t(first + i) * uint64_t(max_tries) + uint64_t(t)) * SQMP_DOF;
#pragma unroll
    for (int j = 0; j < SQMP_DOF; ++j) {
      const double u = __dmul_rn(double(splitmix_at(seed, k0 + j) >> 11), 1.0 / 9007199254740992.0);
      q[j] = __dadd_rn(box.lo[j], __dmul_rn(__dsub_rn(box.hi[j], box.lo[j]), u));
    }
    fk_one<double, SIG>(rb, q, o);
    if (0.0 <= o[2]) {   // above the platform (CL:1182-1184)
      used = t + 1;
      break;
    }
  }
#pragma unroll
  for (int j = 0; j < SQMP_DOF; ++j) q_out[i * SQMP_DOF + j] = q[j];
  if (tries_out) tries_out[i] = used;
}

// sampleJointConfigfromEllipse_JntArray (BS:3607-3829), one thread per sample; same counter stream as k5_sample with nine
// outputs per try (eight joint draws + the scale).  Arithmetic order documented in include/sqmp_b200.h.
template <typename SIG>
__global__ void __launch_bounds__(128)
k5_sample_informed(const __grid_constant__ RobotDev rb, const __grid_constant__ SampleBox box, const __grid_constant__ sqmp_informed el,
                   uint64_t seed, int64_t first, int64_t n, int max_tries, double *__restrict__ q_out,
                   int32_t *__restrict__ tries_out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double q[SQMP_DOF], o[7];
  int used = -max_tries;
  const bool unconfined = el.env_x[0] == 0.0 && el.env_x[1] == 0.0 && el.env_y[0] == 0.0 && el.env_y[1] == 0.0;
  for (int t = 0; t < max_tries; ++t) {
    const uint64_t k0 = (uint64_t(first + i) * uint64_t(max_tries) + uint64_t(t)) * 9ull;
    double ball[SQMP_DOF];
#pragma unroll
    for (int j = 0; j < SQMP_DOF; ++j) {
      const double u = __dmul_rn(double(splitmix_at(seed, k0 + j) >> 11), 1.0 / 9007199254740992.0);
      ball[j] = __dadd_rn(box.lo[j], __dmul_rn(__dsub_rn(box.hi[j], box.lo[j]), u));
    }
    const double scale = __dmul_rn(double(splitmix_at(seed, k0 + 8) >> 11), 1.0 / 9007199254740992.0);
    int out_idx = 0;
    for (int cls = 0; cls < 2; ++cls) {   // prismatic joints first, then revolute ones (BS:3773-3785)
      const int nc = cls == 0 ? el.n_pri : el.n_rev;
      const int *idx = cls == 0 ? el.pri_idx : el.rev_idx;
      const double *Cm = cls == 0 ? el.c_pri : el.c_rev;
      const double *cen = cls == 0 ? el.center_pri : el.center_rev;
      const double *L = cls == 0 ? el.l_pri : el.l_rev;
      double sum = 0.0, y[SQMP_DOF];
      for (int a = 0; a < nc; ++a) sum = __dadd_rn(sum, __dmul_rn(ball[idx[a]], ball[idx[a]]));
      const double norm = sqrt(sum);
      for (int a = 0; a < nc; ++a)   // L (scale * v / |v|)
        y[a] = __dmul_rn(a == 0 ? L[0] : L[1], __dmul_rn(scale, __ddiv_rn(ball[idx[a]], norm)));
      for (int r = 0; r < nc; ++r) {
        double acc = 0.0;
        for (int a = 0; a < nc; ++a) acc = __dadd_rn(acc, __dmul_rn(Cm[r * nc + a], y[a]));
        q[out_idx++] = __dadd_rn(acc, cen[r]);
      }
    }
    fk_one<double, SIG>(rb, q, o);
    const bool inside = unconfined || (q[0] < el.env_x[1] && q[0] > el.env_x[0] && q[1] < el.env_y[1] && q[1] > el.env_y[0]);
    if (0.0 <= o[2] && inside) {
      used = t + 1;
      break;
    }
  }
#pragma unroll
  for (int j = 0; j < SQMP_DOF; ++j) q_out[i * SQMP_DOF + j] = q[j];
  if (tries_out) tries_out[i] = used;
}

cudaError_t launch_sample_informed(const RobotDev &rb, uint64_t seed, int64_t first, int64_t n, const double *lo, const double *hi,
                                   int max_tries, const sqmp_informed &el, double *d_q_out, int32_t *d_tries_out,
                                   bool allow_specialized, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  SampleBox box;
  for (int j = 0; j < SQMP_DOF; ++j) {
    box.lo[j] = lo[j];
    box.hi[j] = hi[j];
  }
  const unsigned blocks = unsigned((n + 127) / 128);
  if (allow_specialized && chain_matches<SigRobotino>(rb))
    k5_sample_informed<SigRobotino><<<blocks, 128, 0, st>>>(rb, box, el, seed, first, n, max_tries, d_q_out, d_tries_out);
  else
    k5_sample_informed<SigGeneric><<<blocks, 128, 0, st>>>(rb, box, el, seed, first, n, max_tries, d_q_out, d_tries_out);
  return cudaGetLastError();
}

cudaError_t launch_sample(const RobotDev &rb, uint64_t seed, int64_t first, int64_t n, const double *lo, const double *hi,
                          int max_tries, double *d_q_out, int32_t *d_tries_out, bool allow_specialized, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  SampleBox box;
  for (int j = 0; j < SQMP_DOF; ++j) {
    box.lo[j] = lo[j];
    box.hi[j] = hi[j];
  }
  const unsigned blocks = unsigned((n + 127) / 128);
  if (allow_specialized && chain_matches<SigRobotino>(rb))
    k5_sample<SigRobotino><<<blocks, 128, 0, st>>>(rb, box, seed, first, n, max_tries, d_q_out, d_tries_out);
  else
    k5_sample<SigGeneric><<<blocks, 128, 0, st>>>(rb, box, seed, first, n, max_tries, d_q_out, d_tries_out);
  return cudaGetLastError();
}

}  // namespace sqmp
