// FFMA vs FFMA2 issue throughput on sm_100a: N independent accumulator chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fma2b(float a, float2 b, float2 c) {   // broadcast scalar operand
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
template <int MODE, int ILP>
__global__ void k(float *out, int iters, float a, float b) {
  float2 acc[ILP];
  for (int i = 0; i < ILP; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a, b); acc[i].y = fmaf(acc[i].y, a, b); }
      else if (MODE == 1) acc[i] = fma2(acc[i], a2, b2);
      else acc[i] = fma2b(a, acc[i], b2);
    }
  }
  float s = 0;
  for (int i = 0; i < ILP; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int ILP>
void run(const char *name, int warps_per_sm) {
  float *d; cudaMalloc(&d, 148 * 1024 * 4);
  const int threads = 256, blocks = 148 * warps_per_sm * 32 / threads, iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, ILP><<<blocks, threads>>>(d, 100, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<MODE, ILP><<<blocks, threads>>>(d, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fmas = double(blocks) * threads * iters * ILP * 2.0;
  printf("%-28s warps/SM %2d ILP %d: %.1f GFMA/s = %.1f FMA/clk/SM (at 1.965 GHz)\n", name, warps_per_sm, ILP, fmas / ms / 1e6,
         fmas / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(d);
}
int main() {
  for (int w : {16, 32, 64}) {
    run<0, 8>("FFMA (scalar)", w);
    run<1, 8>("FFMA2 (3 packed regs)", w);
    run<2, 8>("FFMA2 (broadcast operand)", w);
  }
  run<0, 4>("FFMA (scalar)", 32); run<1, 4>("FFMA2", 32); run<2, 4>("FFMA2 bcast", 32);
  return 0;
}
