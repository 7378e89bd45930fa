// tools/tc_probe_f16.cu — bring-up check of csrc/tc_umma.cuh on the GPU: one 128 x 128 tile of D = |n|^2 - 2 q.n through TWO
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

__global__ void __launch_bounds__(128, 1) k_probe(const double *q, const double *n, float *D) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4 *sA = reinterpret_cast<float4 *>(smem), *sB = reinterpret_cast<float4 *>(smem + TC_TILE_BYTES);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 2 * TC_TILE_BYTES);
  uint32_t *tptr = reinterpret_cast<uint32_t *>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  {
    double v[8], w[8];
    for (int d = 0; d < 8; ++d) { v[d] = q[tid * 8 + d]; w[d] = n[tid * 8 + d]; }
    tc_write_query_row(sA, tid, v, true);
    tc_write_node_row(sB, tid, w);
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
    tc_mma_f16(tmem, tc_desc(a, TC_CHUNK_BYTES), tc_desc(b, 0), false);                               // [qh | qm] . [-2nh | -2nh]
    tc_mma_f16(tmem, tc_desc(a, 2 * TC_CHUNK_BYTES), tc_desc(b + TC_CHUNK_BYTES, TC_CHUNK_BYTES), true);   // [qh | 1 1 1] . [-2nm | norm]
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
  const int smem = 2 * TC_TILE_BYTES + 64;
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
