// tc_umma.cuh — the few tcgen05 / TMEM / mbarrier primitives the tensor-core near-neighbour scan (nn.cu: k_nn_tc) uses,
// and the operand image both sides of its product are stored in.
//
// The product: D[query][node] = |n|^2 - 2 q.n for a tile of 128 queries against a tile of 128 nodes, as ONE fp32
// accumulator of two tcgen05.mma (kind::f16, M = 128, N = 128, K = 16 each) over operands split into fp16 pieces:
//     q = qh + qm (+ eq),  n = nh + nm (+ en),  |n|^2 = p0 + p1 + p2       (every piece has an 11-bit significand)
//     MMA 1:  [ qh | qm ] . [ -2nh | -2nh ]          MMA 2:  [ qh | 1 1 1 0.. ] . [ -2nm | p0 p1 p2 0.. ]
// Products of two 11-bit significands are exact in the fp32 accumulator's 24 bits; what is lost is qm.nm, the residuals
// eq, en (<= 2^-22 relative, or half an fp16 subnormal step for tiny values) and the accumulation roundings: nn.cu budgets
// 2^-14 R^2 + 2^-19 R + 2^-22 for the lot.  fp16 overflows at 65504: the scan is only used while |n|^2 < 6e4 for the whole tree,
// and a query beyond that gets a zero row and an infinite threshold (every pair of it is evaluated exactly).
//
// Operand image ("K-major, no swizzle" canonical layout of the UMMA shared-memory descriptor): an operand tile is 128 rows
// of three 16-byte K chunks (8 fp16 each); chunk c of row r lives at byte  c * 2048 + r * 16.  Eight consecutive rows of
// one chunk form the 8 x 16 B core matrix (128 contiguous bytes), so the stride between 8-row groups (SBO) is 128 B, and
// the descriptor's leading-dimension offset (LBO) is the distance between the two chunks one K = 16 MMA reads:
//     A, MMA 1: chunks (0, 1), LBO 2048        B, MMA 1: chunks (0, 0), LBO 0  — -2nh is stored once and read twice
//     A, MMA 2: chunks (0, 2), LBO 4096        B, MMA 2: chunks (1, 2), LBO 2048
// The tree keeps this image per 128-node tile in global memory (6 KB, contiguous), so a tile reaches shared memory as ONE
// 1-D bulk copy (cp.async.bulk, SASS UBLKCP) — no tensor map.  (tools/tc_probe_f16.cu checks the arithmetic on the GPU.)
#pragma once
#include <cstdint>
#include <cuda_fp16.h>

namespace sqmp {

constexpr int TC_ROWS = 128;                    // rows of an operand tile: queries of a query tile / nodes of a node tile
constexpr int TC_CHUNKS = 3;                    // 16-byte K chunks per row
constexpr int TC_CHUNK_BYTES = TC_ROWS * 16;    // 2048
constexpr int TC_TILE_BYTES = TC_CHUNKS * TC_CHUNK_BYTES;   // 6144
constexpr int TC_TILE_FLOAT4 = TC_TILE_BYTES / 16;
constexpr double TC_MAX_NORM2 = 6.0e4;          // |x|^2 below this: every coordinate and the norm itself fit fp16

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

struct alignas(16) TcChunk {   // eight fp16 = one 16-byte K chunk of one row
  __half v[8];
};
// x = h + m + e with fp16 h, m: |e| <= 2^-22 |x|, or half a subnormal step (2^-25) when m is tiny
__device__ __forceinline__ void fp16_split2(double x, __half &h, __half &m) {
  h = __double2half(x);
  m = __double2half(x - double(__half2float(h)));
}

// image row of one node (B operand): chunk 0 = -2 nh, 1 = -2 nm, 2 = (p0, p1, p2, 0..)
__device__ __forceinline__ void tc_write_node_row(float4 *tile, int r, const double (&v)[8]) {
  TcChunk hi, mid, nrm;
  double n2 = 0.0;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    __half h, m;
    fp16_split2(v[d], h, m);
    hi.v[d] = __float2half(-2.f * __half2float(h));     // exact (a power of two)
    mid.v[d] = __float2half(-2.f * __half2float(m));
    n2 += v[d] * v[d];
    nrm.v[d] = __float2half(0.f);
  }
  nrm.v[0] = __double2half(n2);
  double rest = n2 - double(__half2float(nrm.v[0]));
  nrm.v[1] = __double2half(rest);
  rest -= double(__half2float(nrm.v[1]));
  nrm.v[2] = __double2half(rest);
  TcChunk *t = reinterpret_cast<TcChunk *>(tile);
  t[0 * TC_ROWS + r] = hi;
  t[1 * TC_ROWS + r] = mid;
  t[2 * TC_ROWS + r] = nrm;
}
// image row of one query (A operand): chunk 0 = qh, 1 = qm, 2 = (1, 1, 1, 0..); usable = false: a zero row
__device__ __forceinline__ void tc_write_query_row(float4 *tile, int r, const double (&v)[8], bool usable) {
  TcChunk hi, mid, one;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    fp16_split2(usable ? v[d] : 0.0, hi.v[d], mid.v[d]);
    one.v[d] = __float2half(d < 3 ? 1.f : 0.f);
  }
  TcChunk *t = reinterpret_cast<TcChunk *>(tile);
  t[0 * TC_ROWS + r] = hi;
  t[1 * TC_ROWS + r] = mid;
  t[2 * TC_ROWS + r] = one;
}

// ---- shared-memory matrix descriptor (64 bit): start >> 4 [0,14) | LBO >> 4 [16,30) | SBO >> 4 [32,46) | version 1 [46,48)
//      | base offset 0 [49,52) | layout type 0 = no swizzle [61,64)
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  return uint64_t((smem_addr & 0x3ffffu) >> 4) | (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(128 >> 4) << 32) | (uint64_t(1) << 46);
}
// ---- instruction descriptor (32 bit): D fp32 [4,6) = 1 | A fp16 [7,10) = 0 | B fp16 [10,13) = 0 | A, B K-major [15], [16] = 0
//      | N >> 3 [17,23) | M >> 4 [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (uint32_t(TC_ROWS >> 3) << 17) | (uint32_t(TC_ROWS >> 4) << 24);

// D[tmem] (+)= A[smem] . B[smem]^T; one thread issues it for the CTA
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
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
