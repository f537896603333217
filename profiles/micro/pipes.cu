// Micro-benchmark: per-SM throughput of MUFU.EX2, FFMA, and a degree-3 polynomial exp2 on the FMA/ALU pipes,
// alone and mixed, at 1 / 2 / 4 warps per SMSP. Build: nvcc -arch=sm_100a -O3 pipes.cu -o pipes
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float poly2(float x) {
  x = fmaxf(x, -126.f);
  const float r = x + 12582912.f;
  const float n = r - 12582912.f;
  const float f = x - n;
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t cvt_h2(float lo, float hi) { uint32_t y; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed * (i + 1) + threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]);                       // MUFU only
      if (MODE == 1) a[i] = fmaf(a[i], 0.999f, 0.001f);      // FFMA only
      if (MODE == 2) a[i] = poly2(a[i]);                     // polynomial exp2
      if (MODE == 3) a[i] = ex2(fmaf(a[i], 0.999f, -0.5f)) + 0.25f;  // FFMA + MUFU + FADD (softmax-like)
      if (MODE == 4) a[i] = (i & 3) == 3 ? poly2(fmaf(a[i], 0.999f, -0.5f)) + 0.25f : ex2(fmaf(a[i], 0.999f, -0.5f)) + 0.25f;
      if (MODE == 6) { uint32_t h = ex2_h2(__float_as_uint(a[i])); a[i] = __uint_as_float(h ^ 0x00010001u); }   // 2 elements per op
      if (MODE == 7 && (i & 1) == 0) {  // softmax-like on a PAIR: 2 FFMA (or FFMA2) + cvt.f16x2 + ex2.f16x2 + hadd2
        uint32_t h = ex2_h2(cvt_h2(fmaf(a[i], 0.999f, -0.5f), fmaf(a[i + 1], 0.999f, -0.5f)));
        __half2 hh = *reinterpret_cast<__half2*>(&h);
        hh = __hadd2(hh, __float2half2_rn(0.25f));
        float2 f = __half22float2(hh);
        a[i] = f.x; a[i + 1] = f.y;
      }
      if (MODE == 5) a[i] = (i & 1) ? poly2(fmaf(a[i], 0.999f, -0.5f)) + 0.25f : ex2(fmaf(a[i], 0.999f, -0.5f)) + 0.25f;
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int threads) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, threads>>>(out, 16, -0.3f);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(out, iters, -0.3f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double elems = double(iters) * 16 * threads * (MODE == 6 ? 2 : 1);  // per SM (f16x2: two elements per op)
  double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-34s threads/SM %4d: %.2f elem/clk/SM (%.3f ms)\n", name, threads, elems / cycles, ms);
  cudaFree(out);
}
int main() {
  for (int th : {128, 256, 512}) {
    run<0>("MUFU.EX2", th);
    run<1>("FFMA", th);
    run<2>("poly exp2 (FMA+ALU pipes)", th);
    run<3>("FFMA+MUFU+FADD", th);
    run<4>("softmax-like, 1/4 poly", th);
    run<5>("softmax-like, 1/2 poly", th);
    run<6>("MUFU.EX2 f16x2 (count x2)", th);
    run<7>("softmax-like via f16x2 (per element)", th);
  }
  return 0;
}
