// tc_umma.cuh — the few tcgen05 / TMEM / mbarrier primitives the tensor-core near-neighbour scan (nn.cu: k_nn_tc) uses,
// and the operand image both sides of its product are stored in.
//
// The product: D[query][node] = |n|^2 - 2 q.n for a tile of 128 queries against a tile of 128 nodes, as ONE accumulator
// of four tcgen05.mma (kind::tf32, M = 128, N = 128, K = 8 each) over operands split into tf32 pieces:
//     q = qh + qm (+ eq),  n = nh + nm (+ en),  |n|^2 = p0 + p1 + p2       (every piece has an 11-bit significand)
//     D = qh.(-2nh) + qm.(-2nh) + qh.(-2nm) + (1,1,1).(p0,p1,p2)
// Products of two 11-bit significands are exact in the fp32 accumulator's 24 bits; what is lost is qm.nm and the
// residuals eq, en (each <= 2^-22 relative) and the accumulation roundings: nn.cu budgets 2^-14 R^2 for the lot.
//
// Operand image ("K-major, no swizzle" canonical layout of the UMMA shared-memory descriptor): an operand tile is 128 rows
// of K = 24 tf32 values kept as six 16-byte K chunks; chunk c of row r lives at byte  c * 2048 + r * 16.  Eight
// consecutive rows of one chunk form the 8 x 16 B core matrix (128 contiguous bytes), so
//     stride between 8-row groups (SBO)      = 128 B
//     stride between the K chunks  (LBO)     = 2048 B
// and one MMA (K = 8 = two chunks) addresses chunks (c, c + 1) by starting its descriptor at c * 2048.
// The tree keeps this image per 128-node tile in global memory (12 KB, contiguous), so a tile reaches shared memory as ONE
// 1-D bulk copy (cp.async.bulk, SASS UBLKCP) — no tensor map.
#pragma once
#include <cstdint>

namespace sqmp {

constexpr int TC_ROWS = 128;                    // rows of an operand tile: queries of a query tile / nodes of a node tile
constexpr int TC_CHUNKS = 6;                    // 16-byte K chunks per row
constexpr int TC_CHUNK_BYTES = TC_ROWS * 16;    // 2048
constexpr int TC_TILE_BYTES = TC_CHUNKS * TC_CHUNK_BYTES;   // 12288
constexpr int TC_TILE_FLOAT4 = TC_TILE_BYTES / 16;

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

// round to tf32 (nearest, ties away), low 13 bits cleared: the value the tensor core multiplies
__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r & 0xffffe000u);
}
// x = h + m + e with tf32 h, m and |e| <= 2^-22 |x|   (x - h is exact in fp32)
__device__ __forceinline__ void tf32_split2(float x, float &h, float &m) {
  h = tf32_round(x);
  m = tf32_round(x - h);
}
// three tf32 pieces of a non-negative double (|residual| <= 2^-33 v)
__device__ __forceinline__ void tf32_split3(double v, float &p0, float &p1, float &p2) {
  p0 = tf32_round(__double2float_rn(v));
  const double r1 = v - double(p0);
  p1 = tf32_round(__double2float_rn(r1));
  const double r2 = r1 - double(p1);
  p2 = tf32_round(__double2float_rn(r2));
}

// image row of one node (B operand): chunks 0,1 = -2 nh; 2,3 = -2 nm; 4 = (p0, p1, p2, 0); 5 = 0
__device__ __forceinline__ void tc_write_node_row(float4 *tile, int r, const double (&v)[8]) {
  float h[8], m[8];
  double n2 = 0.0;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    tf32_split2(__double2float_rn(v[d]), h[d], m[d]);
    n2 += v[d] * v[d];
  }
  float p0, p1, p2;
  tf32_split3(n2, p0, p1, p2);
  tile[0 * TC_ROWS + r] = make_float4(-2.f * h[0], -2.f * h[1], -2.f * h[2], -2.f * h[3]);
  tile[1 * TC_ROWS + r] = make_float4(-2.f * h[4], -2.f * h[5], -2.f * h[6], -2.f * h[7]);
  tile[2 * TC_ROWS + r] = make_float4(-2.f * m[0], -2.f * m[1], -2.f * m[2], -2.f * m[3]);
  tile[3 * TC_ROWS + r] = make_float4(-2.f * m[4], -2.f * m[5], -2.f * m[6], -2.f * m[7]);
  tile[4 * TC_ROWS + r] = make_float4(p0, p1, p2, 0.f);
  tile[5 * TC_ROWS + r] = make_float4(0.f, 0.f, 0.f, 0.f);
}
// image row of one query (A operand): chunks 0,1 = qh; 2,3 = qm; 4 = (1, 1, 1, 0); 5 = 0
__device__ __forceinline__ void tc_write_query_row(float4 *tile, int r, const double (&v)[8]) {
  float h[8], m[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) tf32_split2(__double2float_rn(v[d]), h[d], m[d]);
  tile[0 * TC_ROWS + r] = make_float4(h[0], h[1], h[2], h[3]);
  tile[1 * TC_ROWS + r] = make_float4(h[4], h[5], h[6], h[7]);
  tile[2 * TC_ROWS + r] = make_float4(m[0], m[1], m[2], m[3]);
  tile[3 * TC_ROWS + r] = make_float4(m[4], m[5], m[6], m[7]);
  tile[4 * TC_ROWS + r] = make_float4(1.f, 1.f, 1.f, 0.f);
  tile[5 * TC_ROWS + r] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- shared-memory matrix descriptor (64 bit): start >> 4 [0,14) | LBO >> 4 [16,30) | SBO >> 4 [32,46) | version 1 [46,48)
//      | base offset 0 [49,52) | layout type 0 = no swizzle [61,64)
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr) {
  return uint64_t((smem_addr & 0x3ffffu) >> 4) | (uint64_t(TC_CHUNK_BYTES >> 4) << 16) | (uint64_t(128 >> 4) << 32) | (uint64_t(1) << 46);
}
// ---- instruction descriptor (32 bit): D fp32 [4,6) = 1 | A tf32 [7,10) = 2 | B tf32 [10,13) = 2 | A, B K-major [15], [16] = 0
//      | N >> 3 [17,23) | M >> 4 [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(TC_ROWS >> 3) << 17) | (uint32_t(TC_ROWS >> 4) << 24);

// D[tmem] (+)= A[smem] . B[smem]^T; one thread issues it for the CTA
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(TC_IDESC), "r"(uint32_t(accumulate))
      : "memory");
}
// the mbarrier gets one arrival when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMEM: the whole warp calls these
__device__ __forceinline__ void tc_tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_tmem_free(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// 32 consecutive fp32 columns of the calling lane's TMEM row (lane i of warp w reads TMEM lane 32 (w % 4) + i)
__device__ __forceinline__ void tc_tmem_ld32(uint32_t taddr, float *v) {   // v: 32 registers after unrolling
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- mbarriers
__device__ __forceinline__ void tc_mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TC_DONE;\n"
      "bra TC_WAIT;\n"
      "TC_DONE:\n"
      "}\n" ::"r"(tc_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(tc_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ float tc_min3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

}  // namespace sqmp
