// micro-benchmark: per-SM throughput of ex2.approx, bf16 pack (F2FP), FFMA2 and mixes, many warps, no memory traffic
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) v[i] = ex2(v[i]) - 1.0f;                                    // MUFU + FADD
      if (MODE == 1) { __nv_bfloat162 p = __floats2bfloat162_rn(v[i], v[(i + 1) & 15]); acc += *reinterpret_cast<unsigned*>(&p); v[i] += 1.0f; }  // F2FP + IADD + FADD
      if (MODE == 2) { float e = ex2(v[i]); __nv_bfloat162 p = __floats2bfloat162_rn(e, v[(i + 1) & 15]); acc += *reinterpret_cast<unsigned*>(&p); v[i] = e - 1.0f; }  // MUFU + F2FP
      if (MODE == 3) v[i] = fmaf(v[i], 1.0001f, 0.5f);                            // FFMA only
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
}
template <int MODE>
void run(const char* name, int warps_per_sm) {
  int iters = 4096;
  int threads = 256, blocks = 148 * (warps_per_sm * 32 / threads);
  float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<blocks, threads>>>(out, 16, 0.5f);
  cudaEventRecord(a);
  k<MODE><<<blocks, threads>>>(out, iters, 0.5f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double ops = (double)blocks * threads * iters * 16;
  printf("%-28s warps/SM %2d: %.3f ms  %.1f Gop/s  = %.2f lane-ops/clk/SM @ %.0f MHz (nominal max clock)\n", name, warps_per_sm, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1e3);
  cudaFree(out);
}
int main() {
  for (int w : {8, 16, 32}) {
    run<0>("ex2+fadd", w);
    run<1>("f2fp.bf16x2+iadd+fadd", w);
    run<2>("ex2+f2fp+iadd+fadd", w);
    run<3>("ffma", w);
  }
  return 0;
}
