// tools/tc_probe_f16.cu — bring-up check: the same 128 x 128 tile of D = |n|^2 - 2 q.n as tc_probe.cu, through TWO
// tcgen05.mma kind::f16 (K = 16) over fp16-split operands, with the B descriptor's leading-dimension offset set to 0 so
// that both 16-byte K chunks of the first MMA read the same -2 nh.  Development aid, not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I squirrel_motion_planner_b200/csrc -o build/tc_probe_f16 tools/tc_probe_f16.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>

#include "tc_umma.cuh"

using namespace sqmp;

constexpr int CH = 128 * 16;   // bytes of one 16-byte K chunk over 128 rows
__device__ __forceinline__ uint64_t desc16(uint32_t addr, uint32_t lbo_bytes) {
  return uint64_t((addr & 0x3ffffu) >> 4) | (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(128 >> 4) << 32) | (uint64_t(1) << 46);
}
// D fp32 | A, B fp16 (format 0) | K-major | N = 128 | M = 128
constexpr uint32_t IDESC16 = (1u << 4) | (uint32_t(128 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
__device__ __forceinline__ void mma16(uint32_t d, uint64_t a, uint64_t b, bool acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(IDESC16), "r"(uint32_t(acc))
      : "memory");
}
struct H8 { __half v[8]; };   // one 16-byte chunk
__device__ void split2(double x, __half &h, __half &m) {
  h = __double2half(x);
  m = __double2half(x - double(__half2float(h)));
}

__global__ void __launch_bounds__(128, 1) k_probe(const double *q, const double *n, float *D) {
  extern __shared__ __align__(128) unsigned char smem[];
  H8 *sA = reinterpret_cast<H8 *>(smem);            // chunks: 0 qh | 1 qm | 2 ones
  H8 *sB = reinterpret_cast<H8 *>(smem + 3 * CH);   // chunks: 0 -2nh | 1 -2nm | 2 norm pieces
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 6 * CH);
  uint32_t *tptr = reinterpret_cast<uint32_t *>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  {
    H8 qh, qm, one, nh, nm, nr;
    double n2 = 0.0;
    for (int d = 0; d < 8; ++d) {
      split2(q[tid * 8 + d], qh.v[d], qm.v[d]);
      __half h, m;
      split2(n[tid * 8 + d], h, m);
      nh.v[d] = __float2half(-2.f * __half2float(h));
      nm.v[d] = __float2half(-2.f * __half2float(m));
      n2 += n[tid * 8 + d] * n[tid * 8 + d];
      one.v[d] = __float2half(d < 3 ? 1.f : 0.f);
      nr.v[d] = __float2half(0.f);
    }
    nr.v[0] = __double2half(n2);
    double r = n2 - double(__half2float(nr.v[0]));
    nr.v[1] = __double2half(r);
    r -= double(__half2float(nr.v[1]));
    nr.v[2] = __double2half(r);
    sA[0 * 128 + tid] = qh; sA[1 * 128 + tid] = qm; sA[2 * 128 + tid] = one;
    sB[0 * 128 + tid] = nh; sB[1 * 128 + tid] = nm; sB[2 * 128 + tid] = nr;
  }
  tc_fence_async_smem();
  if (tid == 0) { tc_mbar_init(bar, 1); tc_fence_mbar_init(); }
  if (warp == 0) tc_tmem_alloc(tptr, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (tid == 0) {
    const uint32_t a = tc_smem_u32(sA), b = tc_smem_u32(sB);
    mma16(tmem, desc16(a, CH), desc16(b, 0), false);               // [qh | qm] . [-2nh | -2nh]
    mma16(tmem, desc16(a, 2 * CH), desc16(b + CH, CH), true);      // [qh | 1 1 1] . [-2nm | norm]
    tc_commit(bar);
  }
  tc_mbar_wait(bar, 0);
  tc_fence_after();
  for (int c = 0; c < 128; c += 32) {
    float v[32];
    tc_tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + c, v);
    tc_tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[tid * 128 + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_tmem_free(tmem, 128);
}

int main() {
  std::vector<double> q(128 * 8), n(128 * 8);
  srand(7);
  for (auto &x : q) x = (rand() / double(RAND_MAX) - 0.5) * 10.0;
  for (auto &x : n) x = (rand() / double(RAND_MAX) - 0.5) * 10.0;
  double *dq, *dn;
  float *dD;
  cudaMalloc(&dq, q.size() * 8); cudaMalloc(&dn, n.size() * 8); cudaMalloc(&dD, 128 * 128 * 4);
  cudaMemcpy(dq, q.data(), q.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dn, n.data(), n.size() * 8, cudaMemcpyHostToDevice);
  const int smem = 6 * CH + 64;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_probe<<<1, 128, smem>>>(dq, dn, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> D(128 * 128);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0.0, R2 = 0.0;
  for (int i = 0; i < 128; ++i) {
    double a = 0, b = 0;
    for (int d = 0; d < 8; ++d) { a += q[i * 8 + d] * q[i * 8 + d]; b += n[i * 8 + d] * n[i * 8 + d]; }
    R2 = fmax(R2, fmax(a, b));
  }
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j) {
      double nn = 0, dot = 0;
      for (int d = 0; d < 8; ++d) { nn += n[j * 8 + d] * n[j * 8 + d]; dot += q[i * 8 + d] * n[j * 8 + d]; }
      worst = fmax(worst, fabs(double(D[i * 128 + j]) - (nn - 2.0 * dot)));
    }
  printf("tc_probe_f16: max |D - (|n|^2 - 2 q.n)| = %.3e = 2^%.1f R^2 (R^2 = %.1f); D[0][0] = %.6f D[5][77] = %.6f\n", worst,
         log2(worst / R2 + 1e-300), R2, D[0], D[5 * 128 + 77]);
  printf(worst < R2 / 16384.0 ? "tc_probe_f16: OK\n" : "tc_probe_f16: FAIL\n");
  return 0;
}
