// Micro-benchmark: cycles per 32x128 softmax block of ONE warp (the per-KV-block register work of the attention
// kernel's softmax warps, without TMEM / MMA / barriers), at 1, 2 and 3 warps per SM sub-partition. Separates the cost
// of each ingredient (row max, scale, MUFU.EX2, row sum, bf16 pack, polynomial exp2 share) so that the per-block
// budget of attn_fwd3_kernel can be set against what the pipes deliver.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 softmax_pipe.cu -o softmax_pipe
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float max3f(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)), "l"(reinterpret_cast<const uint64_t&>(c)));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float2 fadd2_rm(float2 a, float2 b) {
  float2 d;
  asm("add.rm.ftz.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  const float kMagic = 12582912.f;
  x.x = fmaxf(x.x, -127.f);
  x.y = fmaxf(x.y, -127.f);
  const float2 t = fadd2_rm(x, make_float2(kMagic, kMagic));
  const float2 n = fadd2(t, make_float2(-kMagic, -kMagic));
  const float2 f = ffma2(n, make_float2(-1.f, -1.f), x);
  float2 p = ffma2(f, make_float2(0.077119089663028717f, 0.077119089663028717f), make_float2(0.227564394474029541f, 0.227564394474029541f));
  p = ffma2(p, f, make_float2(0.695146143436431885f, 0.695146143436431885f));
  p = ffma2(p, f, make_float2(1.f, 1.f));
  float2 r;
  r.x = __int_as_float((__float_as_int(t.x) << 23) + __float_as_int(p.x));
  r.y = __int_as_float((__float_as_int(t.y) << 23) + __float_as_int(p.y));
  return r;
}

// bit flags
enum { F_MAX = 1, F_SCALE = 2, F_EXP = 4, F_SUM = 8, F_PACK = 16, F_TMEM = 32 };

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

template <int FLAGS, int POLY8>
__global__ void __launch_bounds__(384, 1) k(float* out, long long* cyc, int iters, float seed) {
  __shared__ uint32_t tmem_slot;
  uint32_t tbase = 0;
  if (FLAGS & F_TMEM) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w: lane quadrant w % 4; warps of the same quadrant use different column ranges
    tbase = tmem_slot + ((uint32_t)((threadIdx.x >> 5) & 3) * 32u << 16) + (threadIdx.x >> 7) * 160;
  }
  float2 x2[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) x2[i] = make_float2(seed * (i + 1) + threadIdx.x * 1e-6f, seed * (i + 2));
  float m_ref = 0.f, l = 0.f;
  uint32_t acc = 0;
  const float scale = 0.18f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float* x = reinterpret_cast<float*>(x2);
    // regenerate the block (stands in for the TMEM load): 64 packed adds
    const float2 d = make_float2(-0.37f + 1e-3f * it, 0.21f);
    if (FLAGS & F_TMEM) {
      uint32_t* xr = reinterpret_cast<uint32_t*>(x2);
      tmem_ld32(tbase + 0, xr); tmem_ld32(tbase + 32, xr + 32); tmem_ld32(tbase + 64, xr + 64); tmem_ld32(tbase + 96, xr + 96);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 64; ++i) { x2[i].x = __uint_as_float((__float_as_uint(x2[i].x) & 0x007fffffu) | 0x3f800000u) + d.x; x2[i].y = __uint_as_float((__float_as_uint(x2[i].y) & 0x007fffffu) | 0x3f800000u); }
    } else {
#pragma unroll
      for (int i = 0; i < 64; ++i) x2[i] = fadd2(x2[i], d);
    }
    if (FLAGS & F_MAX) {
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 128; c += 8) {
        mx0 = max3f(mx0, x[c], x[c + 1]);
        mx1 = max3f(mx1, x[c + 2], x[c + 3]);
        mx2 = max3f(mx2, x[c + 4], x[c + 5]);
        mx3 = max3f(mx3, x[c + 6], x[c + 7]);
      }
      const float mx = fmaxf(fmaxf(mx0, mx2), fmaxf(mx1, mx3)) * scale;
      if (mx > m_ref + 8.f) m_ref = mx;
    }
    const float2 neg_m2 = make_float2(-m_ref, -m_ref), scale2 = make_float2(scale, scale);
    float2 a0 = make_float2(0.f, 0.f), a1 = a0;
    uint32_t pk[64];
#pragma unroll
    for (int i = 0; i < 64; i += 2) {
      float2 v0 = x2[i], v1 = x2[i + 1];
      if (FLAGS & F_SCALE) { v0 = ffma2(v0, scale2, neg_m2); v1 = ffma2(v1, scale2, neg_m2); }
      if (FLAGS & F_EXP) {
        if ((i & 7) < POLY8) v0 = ex2_poly2(v0); else v0 = make_float2(ex2(v0.x), ex2(v0.y));
        if (((i + 1) & 7) < POLY8) v1 = ex2_poly2(v1); else v1 = make_float2(ex2(v1.x), ex2(v1.y));
      }
      if (FLAGS & F_SUM) { a0 = fadd2(a0, v0); a1 = fadd2(a1, v1); }
      if (FLAGS & F_PACK) {
        if (FLAGS & F_TMEM) {
          pk[i] = pack_bf16x2(v0.x, v0.y); pk[i + 1] = pack_bf16x2(v1.x, v1.y);
          if ((i & 15) == 14) tmem_st16(tbase + 128 + (i >> 4) * 16 - 0, pk + (i & ~15));
        } else {
          acc ^= pack_bf16x2(v0.x, v0.y) ^ pack_bf16x2(v1.x, v1.y);
        }
      }
      x2[i] = v0; x2[i + 1] = v1;
    }
    l += (a0.x + a0.y) + (a1.x + a1.y);
    if (FLAGS & F_TMEM) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  float s = l + __uint_as_float(acc & 0x3fffffff);
#pragma unroll
  for (int i = 0; i < 64; ++i) s += x2[i].x + x2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  if (FLAGS & F_TMEM) {
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512) : "memory");
  }
}

template <int FLAGS, int POLY8>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 384 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2048;
  printf("%-44s", name);
  for (int threads : {128, 256, 384}) {
    k<FLAGS, POLY8><<<148, threads>>>(out, cyc, 16, -0.003f);
    k<FLAGS, POLY8><<<148, threads>>>(out, cyc, iters, -0.003f);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_block = double(h) / iters;              // cycles per loop iteration of one warp
    printf("  %dw/SMSP: %7.0f cyc/iter (%6.0f per warp-block)", threads / 128, per_block, per_block / (threads / 128));
  }
  printf("\n");
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 0>("regen only (64 FADD2)");
  run<F_MAX, 0>("+ row max (64 FMNMX3)");
  run<F_SCALE, 0>("+ scale (64 FFMA2)");
  run<F_EXP, 0>("+ 128 MUFU.EX2");
  run<F_PACK, 0>("+ pack (64 F2FP + 32 LOP3)");
  run<F_SUM, 0>("+ row sum (64 FADD2)");
  run<F_EXP | F_PACK, 0>("+ MUFU + pack");
  run<F_EXP | F_SUM, 0>("+ MUFU + sum");
  run<F_SCALE | F_EXP | F_SUM | F_PACK, 0>("+ scale, MUFU, sum, pack");
  run<F_MAX | F_SCALE | F_EXP | F_SUM | F_PACK, 0>("full softmax block");
  run<F_MAX | F_SCALE | F_EXP | F_SUM | F_PACK, 1>("full, 1/8 poly");
  run<F_MAX | F_SCALE | F_EXP | F_SUM | F_PACK, 2>("full, 2/8 poly");
  run<F_MAX | F_SCALE | F_EXP | F_SUM | F_PACK, 3>("full, 3/8 poly");
  run<F_MAX | F_SCALE | F_EXP | F_SUM | F_PACK, 4>("full, 4/8 poly");
  run<F_MAX | F_SCALE | F_EXP | F_SUM | F_PACK, 8>("full, all poly");
  run<F_EXP, 8>("+ poly exp2 only (no MUFU)");
  run<F_TMEM, 0>("TMEM: 4 LDTM.x32 + bit fix (no regen)");
  run<F_TMEM | F_PACK, 0>("TMEM: LDTM + pack + 4 STTM.x16");
  run<F_TMEM | F_SCALE | F_EXP | F_SUM | F_PACK, 0>("TMEM: LDTM, scale, MUFU, sum, pack, STTM");
  run<F_TMEM | F_MAX | F_SCALE | F_EXP | F_SUM | F_PACK, 0>("TMEM: full softmax block");
  run<F_TMEM | F_MAX | F_SCALE | F_EXP | F_SUM | F_PACK, 1>("TMEM: full, 1/8 poly");
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
