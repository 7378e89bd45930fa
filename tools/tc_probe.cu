// tools/tc_probe.cu — bring-up check of csrc/tc_umma.cuh on the GPU: one 128 x 128 tile of D = |n|^2 - 2 q.n through
// tcgen05.mma (tf32 split operands) against the same values in fp64.  Prints the largest error in units of R^2 (the
// scan in nn.cu budgets 2^-14 R^2).  Development aid, not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I squirrel_motion_planner_b200/csrc -o build/tc_probe tools/tc_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_umma.cuh"

using namespace sqmp;

__global__ void __launch_bounds__(128, 1) k_probe(const double *q, const double *n, float *D) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4 *sA = reinterpret_cast<float4 *>(smem);
  float4 *sB = reinterpret_cast<float4 *>(smem + TC_TILE_BYTES);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 2 * TC_TILE_BYTES);
  uint32_t *tptr = reinterpret_cast<uint32_t *>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    double v[8], w[8];
    for (int d = 0; d < 8; ++d) { v[d] = q[tid * 8 + d]; w[d] = n[tid * 8 + d]; }
    tc_write_query_row(sA, tid, v);
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
    tc_mma_tf32(tmem, tc_desc(a), tc_desc(b), false);
    tc_mma_tf32(tmem, tc_desc(a + 2 * TC_CHUNK_BYTES), tc_desc(b), true);
    tc_mma_tf32(tmem, tc_desc(a), tc_desc(b + 2 * TC_CHUNK_BYTES), true);
    tc_mma_tf32(tmem, tc_desc(a + 4 * TC_CHUNK_BYTES), tc_desc(b + 4 * TC_CHUNK_BYTES), true);
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
  int bad_i = -1, bad_j = -1;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j) {
      double nn = 0, dot = 0;
      for (int d = 0; d < 8; ++d) { nn += n[j * 8 + d] * n[j * 8 + d]; dot += q[i * 8 + d] * n[j * 8 + d]; }
      const double err = fabs(double(D[i * 128 + j]) - (nn - 2.0 * dot));
      if (err > worst) { worst = err; bad_i = i; bad_j = j; }
    }
  printf("tc_probe: max |D - (|n|^2 - 2 q.n)| = %.3e = 2^%.1f R^2 (R^2 = %.1f) at (%d, %d); D[0][0] = %.6f D[5][77] = %.6f\n", worst,
         log2(worst / R2 + 1e-300), R2, bad_i, bad_j, D[0], D[5 * 128 + 77]);
  printf(worst < R2 / 16384.0 ? "tc_probe: OK (inside the 2^-14 R^2 budget)\n" : "tc_probe: FAIL\n");
  return worst < R2 / 16384.0 ? 0 : 2;
}
