// Throughput probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.   nvcc -arch=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE> __global__ void k(float* out, int iters, float s) {
  float a[16];
  u64 p[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pk2(a[2 * i], a[2 * i + 1]);
  const u64 s2 = pk2(s, s), c2 = pk2(0.5f, 0.25f);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s), "f"(0.5f));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], s2, c2);
    }
  }
  float r = 0;
  if (MODE == 0) { for (int i = 0; i < 16; ++i) r += a[i]; }
  else { for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); r += x + y; } }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 4, 512>>>(out, iters, 0.999f); else k<1><<<148 * 4, 512>>>(out, iters, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fmas = 148.0 * 4 * 512 * 16.0 * iters;
      printf("%s: %.3f ms  %.1f Gfma/s  (%.1f fma/clk/SM at 1.9 GHz)\n", mode ? "FFMA2" : "FFMA ", ms, fmas / ms / 1e6, fmas / ms / 1e6 / 148 / 1.9);
    }
  }
  return 0;
}
