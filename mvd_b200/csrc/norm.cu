// norm.cu — bandwidth-bound normalisation kernels (coalesced 128-bit HBM access, warp-shuffle reductions).
//
//  * GroupNorm(32)(+SiLU) over NHWC bf16, optionally over the channel concatenation of two tensors
//    (diffusers ResnetBlock2D.norm1/norm2, Transformer2DModel.norm, conv_norm_out).
//    Two launches: per-(image, row-chunk) partial sums -> apply (which first reduces the partials).
//  * LayerNorm over the last dim of [M, C] bf16 (BasicTransformerBlock.norm1/2/3) and of small fp32 rows
//    (CameraEncoder MLPs, src/models/camera_encoder.py:31-76).
//  * Reference-feature normalisation of src/models/attention.py:95-103: (r - mean) / clamp(std, 1e-6) * 0.5 with
//    statistics over dims (0,1) of the RAW tensor: per pixel over (batch, channel) for 4-D [B,C,H,W] features,
//    per channel over (batch, sequence) for 3-D [B,S,C] features; unbiased std.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cooperative_groups.h>
#include "host_common.h"
#include "../../include/mvd_b200.h"

namespace mvd {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
// x * sigmoid(x) = h + h * tanh(h), h = x / 2: one MUFU op per element instead of ex2 + a full division. The bf16
// GroupNorm outputs are MUFU/ALU-limited at 64x64 latents otherwise; tanh.approx (rel. error ~2^-11) is far below the
// bf16 rounding of the result.
__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm
// ------------------------------------------------------------------------------------------------
constexpr int GN_THREADS = 256;
constexpr int GN_MAX_VPT = 2;  // 16-byte vectors per thread along C: C <= 256 * 2 * 8 = 4096
constexpr int GN_MAX_GROUPS = 32;
constexpr int GN_MAX_CHUNKS = 128;
constexpr int GN_UNROLL = 8;        // rows in flight per thread (statistics pass)
constexpr int GN_APPLY_UNROLL = 4;  // rows in flight per thread (apply pass; keeps registers < 64)

// chunks (blocks) per image: enough blocks to fill the machine a few times over, at least ~16 rows each
__host__ __device__ inline int gn_chunks(int hw, int n_img, int ty) {
  // rows per chunk ~ ty * GN_UNROLL: each thread needs about one batch of loads
  const int rows = ty * GN_UNROLL;
  int c = (hw + rows - 1) / rows;
  if (c > GN_MAX_CHUNKS) c = GN_MAX_CHUNKS;
  while (c > 1 && static_cast<long>(c) * n_img > 1184) c = (c + 1) / 2;  // <= 8 blocks per SM
  return c < 1 ? 1 : c;
}

// Thread layout shared by both kernels: block = TX x TY threads (padded to whole warps), thread (tx, ty) owns the
// channel vectors tx, tx + TX (VPT of them) of every TY-th row of the block's row chunk: fixed channels per thread
// (per-channel scale/shift and partial sums live in registers), consecutive tx -> consecutive 16-byte vectors
// (coalesced), GN_UNROLL rows in flight per thread.
struct GnSrc {
  const __nv_bfloat16* base;  // first row of this image, at the thread's channel offset
  int cs;                     // row stride (elements) of the source tensor
};
__device__ __forceinline__ GnSrc gn_src(const __nv_bfloat16* x1, int c1, const __nv_bfloat16* x2, int c2, int n, int hw,
                                        int c) {
  GnSrc s;
  if (c < c1) {
    s.base = x1 + static_cast<int64_t>(n) * hw * c1 + c;
    s.cs = c1;
  } else {
    s.base = x2 + static_cast<int64_t>(n) * hw * c2 + (c - c1);
    s.cs = c2;
  }
  return s;
}

// Pass 1. Every block reduces its row chunk to per-group (sum, sumsq) partials (fixed order, no atomics on data);
// the LAST block of an image to finish (ticket counter, self-resetting -> CUDA-graph safe) combines the partials of
// all chunks in chunk order (fp64) into mean / rstd. Result: stats[n][g] = (mean, rstd). Deterministic.
template <int VPT>
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x1, int c1, const __nv_bfloat16* __restrict__ x2, int c2, int hw,
                int groups, int TX, int TY, float eps, float* __restrict__ partial, float* __restrict__ stats,
                unsigned int* __restrict__ tickets) {
  pdl_wait();
  pdl_launch_dependents();
  const int C = c1 + c2;
  const int cpg = C / groups;
  const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int r0 = static_cast<int>(static_cast<int64_t>(hw) * chunk / nchunks);
  const int r1 = static_cast<int>(static_cast<int64_t>(hw) * (chunk + 1) / nchunks);
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const bool active = ty < TY;  // the block is padded to whole warps
  extern __shared__ float s_part[];  // [2][TY][C]
  float* s_psum = s_part;
  float* s_psq = s_part + static_cast<size_t>(TY) * C;
  __shared__ bool s_last;
  {
    GnSrc src[VPT];
    bool on[VPT];
    float sum[VPT][8], sq[VPT][8];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (tx + i * TX) * 8;
      on[i] = active && c < C;
      src[i] = gn_src(x1, c1, x2, c2, n, hw, on[i] ? c : 0);
#pragma unroll
      for (int k = 0; k < 8; ++k) sum[i][k] = sq[i][k] = 0.f;
    }
    // rows in predicated batches of GN_UNROLL, all VPT vectors of a batch together: every load of a batch is issued
    // before any is consumed, also for the last (partial) batch — a serial loop would expose one full memory round
    // trip per row
    for (int r = r0 + ty; r < r1; r += GN_UNROLL * TY) {
      uint4 raw[VPT][GN_UNROLL];
#pragma unroll
      for (int i = 0; i < VPT; ++i)
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u) {
          const int rr = r + u * TY;
          raw[i][u] = (on[i] && rr < r1)
                          ? *reinterpret_cast<const uint4*>(src[i].base + static_cast<int64_t>(rr) * src[i].cs)
                          : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
      for (int i = 0; i < VPT; ++i)
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u) {
          float f[8];
          unpack8(raw[i][u], f);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            sum[i][k] += f[k];
            sq[i][k] += f[k] * f[k];
          }
        }
    }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      if (on[i]) {
        const int c = (tx + i * TX) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          s_psum[static_cast<size_t>(ty) * C + c + k] = sum[i][k];
          s_psq[static_cast<size_t>(ty) * C + c + k] = sq[i][k];
        }
      }
    }
  }
  __syncthreads();
  {
    // 8 threads per group: thread (g, part) adds the channels part, part+8, ... of group g over all TY rows, then a
    // fixed xor-shuffle tree combines the 8 parts (deterministic). blockDim is a multiple of 32; groups <= 32.
    const int part = threadIdx.x & 7;
    for (int g0 = 0; g0 < groups; g0 += blockDim.x >> 3) {
      const int g = g0 + (threadIdx.x >> 3);
      float gs = 0.f, gq = 0.f;
      if (g < groups) {
        for (int cc = part; cc < cpg; cc += 8) {
          const int c = g * cpg + cc;
          for (int y = 0; y < TY; ++y) {
            gs += s_psum[static_cast<size_t>(y) * C + c];
            gq += s_psq[static_cast<size_t>(y) * C + c];
          }
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        gs += __shfl_xor_sync(0xffffffffu, gs, o);
        gq += __shfl_xor_sync(0xffffffffu, gq, o);
      }
      if (g < groups && part == 0) {
        float* p = partial + ((static_cast<int64_t>(n) * nchunks + chunk) * groups + g) * 2;
        p[0] = gs;
        p[1] = gq;
      }
    }
    __threadfence();  // partials visible device-wide before the ticket is taken
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&tickets[n], 1u);
    s_last = (t == static_cast<unsigned int>(nchunks) - 1u);
    if (s_last) tickets[n] = 0u;  // re-arm for the next launch
  }
  __syncthreads();
  if (s_last) {
    // All loads of a slice are independent (issued back to back, L2-coherent __ldcg), then combined in a fixed
    // order: slice-local chunk order, then slice order 0..7 -> deterministic regardless of which block is last.
    __threadfence();
    const int g = threadIdx.x & 31, slice = threadIdx.x >> 5;  // 8 slices x 32 groups (blockDim >= 256 not required)
    const int nslices = blockDim.x >> 5;
    double s = 0.0, q = 0.0;
    if (g < groups) {
      const float* pbase = partial + (static_cast<int64_t>(n) * nchunks * groups + g) * 2;
#pragma unroll 4
      for (int k = slice; k < nchunks; k += nslices) {
        const float2 v = __ldcg(reinterpret_cast<const float2*>(pbase + static_cast<int64_t>(k) * groups * 2));
        s += v.x;
        q += v.y;
      }
    }
    double* red = reinterpret_cast<double*>(s_part);  // reuse: [2][nslices][32] doubles (<= 4 KB)
    __syncthreads();
    red[slice * 32 + g] = s;
    red[(nslices + slice) * 32 + g] = q;
    __syncthreads();
    if (threadIdx.x < groups) {
      double ts = 0.0, tq = 0.0;
      for (int k = 0; k < nslices; ++k) {
        ts += red[k * 32 + threadIdx.x];
        tq += red[(nslices + k) * 32 + threadIdx.x];
      }
      const double cnt = static_cast<double>(hw) * cpg;
      const double mean = ts / cnt;
      double var = tq / cnt - mean * mean;  // biased variance, as torch GroupNorm
      if (var < 0.0) var = 0.0;
      stats[(static_cast<int64_t>(n) * groups + threadIdx.x) * 2] = static_cast<float>(mean);
      stats[(static_cast<int64_t>(n) * groups + threadIdx.x) * 2 + 1] =
          static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
  }
}

// Pass 2: y = (x - mean) * rstd * gamma + beta (+ SiLU), pure streaming.
template <int VPT>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x1, int c1, const __nv_bfloat16* __restrict__ x2, int c2, int hw,
                int groups, int silu, const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                const float* __restrict__ stats, int TX, int TY, __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int C = c1 + c2;
  const int cpg = C / groups;
  const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int r0 = static_cast<int>(static_cast<int64_t>(hw) * chunk / nchunks);
  const int r1 = static_cast<int>(static_cast<int64_t>(hw) * (chunk + 1) / nchunks);
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const bool active = ty < TY;
  GnSrc src[VPT];
  bool on[VPT];
  __nv_bfloat16* obase[VPT];
  float a[VPT][8], b[VPT][8];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (tx + i * TX) * 8;
    on[i] = active && c < C;
    const int cc = on[i] ? c : 0;
    src[i] = gn_src(x1, c1, x2, c2, n, hw, cc);
    obase[i] = out + static_cast<int64_t>(n) * hw * C + cc;
    float gm[8], bt[8];
    unpack8(*reinterpret_cast<const uint4*>(gamma + cc), gm);
    unpack8(*reinterpret_cast<const uint4*>(beta + cc), bt);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int g = (cc + k) / cpg;
      const float mean = stats[(static_cast<int64_t>(n) * groups + g) * 2];
      const float rstd = stats[(static_cast<int64_t>(n) * groups + g) * 2 + 1];
      a[i][k] = gm[k] * rstd;
      b[i][k] = bt[k] - mean * a[i][k];
    }
  }
  for (int r = r0 + ty; r < r1; r += GN_APPLY_UNROLL * TY) {
    uint4 raw[VPT][GN_APPLY_UNROLL];
#pragma unroll
    for (int i = 0; i < VPT; ++i)
#pragma unroll
      for (int u = 0; u < GN_APPLY_UNROLL; ++u) {
        const int rr = r + u * TY;
        raw[i][u] = (on[i] && rr < r1)
                        ? *reinterpret_cast<const uint4*>(src[i].base + static_cast<int64_t>(rr) * src[i].cs)
                        : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
    for (int i = 0; i < VPT; ++i)
#pragma unroll
      for (int u = 0; u < GN_APPLY_UNROLL; ++u) {
        const int rr = r + u * TY;
        float f[8];
        unpack8(raw[i][u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          f[k] = f[k] * a[i][k] + b[i][k];
          if (silu) f[k] = silu_fast(f[k]);
        }
        if (on[i] && rr < r1) *reinterpret_cast<uint4*>(obase[i] + static_cast<int64_t>(rr) * C) = pack8(f);
      }
  }
}

// Single-launch GroupNorm for slabs that fit in shared memory: one block per (image, group). The group's slab
// (hw pixels x cpg channels, bf16) is read from global memory exactly once into shared memory, the exact mean and then
// the centred variance are reduced from there (fixed order -> deterministic), and the normalised, scaled (and SiLU'd)
// values are written straight out. Thread t owns the 32-bit word (channel pair) t % wpp of every ppb-th pixel, so its
// gamma/beta/source pointer are fixed and consecutive threads touch consecutive words.
// (Splitting a slab over a thread-block cluster with a DSMEM exchange of the partial sums was measured slower than
// both this kernel and the two-kernel path at every site: 28 vs 23 us at 8x4096x320.)
constexpr int GN1_THREADS = 512;
constexpr int GN1_UNROLL = 8;
constexpr int GN1_MAX_SMEM = 112 * 1024;  // at least two blocks per SM; larger slabs take the two-kernel path

__device__ __forceinline__ float gn1_block_sum(float v, float* s_red) {
  // xor-shuffle tree per warp, then every thread adds the warp totals in warp order: identical on all threads
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) t += s_red[i];
  return t;
}

__global__ void __launch_bounds__(GN1_THREADS)
gn_fused_kernel(const __nv_bfloat16* __restrict__ x1, int c1, const __nv_bfloat16* __restrict__ x2, int c2, int hw,
                int groups, float eps, int silu, const __nv_bfloat16* __restrict__ gamma,
                const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ uint32_t gn1_slab[];  // [hw][wpp] channel pairs
  __shared__ float s_red[2][GN1_THREADS / 32];
  const int C = c1 + c2;
  const int cpg = C / groups;
  const int wpp = cpg >> 1;  // 32-bit words per pixel
  const int g = blockIdx.x, n = blockIdx.y;
  const int rows = hw;
  const int ppb = blockDim.x / wpp;  // pixels per pass over the block
  const int w = threadIdx.x % wpp, p0 = threadIdx.x / wpp;
  const bool active = p0 < ppb;
  const int c = g * cpg + 2 * w;
  const GnSrc src = gn_src(x1, c1, x2, c2, n, hw, c);
  const int step = ppb * GN1_UNROLL;

  float s = 0.f;
  for (int p = p0; p < rows; p += step) {
    uint32_t raw[GN1_UNROLL];
#pragma unroll
    for (int u = 0; u < GN1_UNROLL; ++u) {
      const int pp = p + u * ppb;
      raw[u] = (active && pp < rows) ? *reinterpret_cast<const uint32_t*>(src.base + static_cast<int64_t>(pp) * src.cs)
                                     : 0u;
    }
#pragma unroll
    for (int u = 0; u < GN1_UNROLL; ++u) {
      const int pp = p + u * ppb;
      if (active && pp < rows) {
        gn1_slab[pp * wpp + w] = raw[u];
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[u]));
        s += f.x + f.y;
      }
    }
  }
  const float cnt = static_cast<float>(hw) * static_cast<float>(cpg);
  const float mean = gn1_block_sum(s, s_red[0]) / cnt;

  // centred second moment from the on-chip copy; a thread re-reads only the words it wrote itself
  float q = 0.f;
  if (active) {
    for (int pp = p0; pp < rows; pp += ppb) {
      const uint32_t r = gn1_slab[pp * wpp + w];
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r));
      const float d0 = f.x - mean, d1 = f.y - mean;
      q += d0 * d0 + d1 * d1;
    }
  }
  const float var = gn1_block_sum(q, s_red[1]) / cnt;  // biased variance, as torch GroupNorm
  const float rstd = rsqrtf(var + eps);

  if (active) {
    const float2 gm = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(gamma + c));
    const float2 bt = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(beta + c));
    const float a0 = gm.x * rstd, a1 = gm.y * rstd;
    const float b0 = bt.x - mean * a0, b1 = bt.y - mean * a1;
    __nv_bfloat16* obase = out + static_cast<int64_t>(n) * hw * C + c;
#pragma unroll 4
    for (int pp = p0; pp < rows; pp += ppb) {
      const uint32_t r = gn1_slab[pp * wpp + w];
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r));
      float y0 = f.x * a0 + b0, y1 = f.y * a1 + b1;
      if (silu) {
        y0 = silu_fast(y0);
        y1 = silu_fast(y1);
      }
      *reinterpret_cast<__nv_bfloat162*>(obase + static_cast<int64_t>(pp) * C) = __floats2bfloat162_rn(y0, y1);
    }
  }
}

// Cluster form of gn_fused_kernel for SMALL batches (a view-sharded rank holds 1-2 samples: 32-64 (image, group) slabs
// for 148 SMs). A cluster of CL CTAs shares one slab by pixel ranges; the two reductions (sum, centred second moment)
// are completed over distributed shared memory: every CTA publishes its partial in its own shared memory, the
// cluster synchronises, every CTA adds the CL partials in rank order (deterministic, identical in all CTAs).
// One launch, CL times the parallelism. (At batch 8 the slabs already fill the machine and this form measured slower.)
__global__ void __launch_bounds__(GN1_THREADS)
gn_fused_cluster_kernel(const __nv_bfloat16* __restrict__ x1, int c1, const __nv_bfloat16* __restrict__ x2, int c2,
                        int hw, int groups, float eps, int silu, const __nv_bfloat16* __restrict__ gamma,
                        const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  extern __shared__ uint32_t gn1_slab[];  // [rows of this CTA][wpp] channel pairs
  __shared__ float s_red[2][GN1_THREADS / 32];
  __shared__ float s_part[2];  // this CTA's partial sum / partial centred second moment (read by the whole cluster)
  const int C = c1 + c2;
  const int cpg = C / groups;
  const int wpp = cpg >> 1;
  const int g = blockIdx.x / CL, n = blockIdx.y;
  const int rows_per = (hw + CL - 1) / CL;
  const int r0 = rank * rows_per;
  const int rows = (r0 + rows_per <= hw) ? rows_per : (hw > r0 ? hw - r0 : 0);
  const int ppb = blockDim.x / wpp;
  const int w = threadIdx.x % wpp, p0 = threadIdx.x / wpp;
  const bool active = p0 < ppb;
  const int c = g * cpg + 2 * w;
  GnSrc src = gn_src(x1, c1, x2, c2, n, hw, c);
  src.base += static_cast<int64_t>(r0) * src.cs;
  const int step = ppb * GN1_UNROLL;

  float s = 0.f;
  for (int p = p0; p < rows; p += step) {
    uint32_t raw[GN1_UNROLL];
#pragma unroll
    for (int u = 0; u < GN1_UNROLL; ++u) {
      const int pp = p + u * ppb;
      raw[u] = (active && pp < rows) ? *reinterpret_cast<const uint32_t*>(src.base + static_cast<int64_t>(pp) * src.cs)
                                     : 0u;
    }
#pragma unroll
    for (int u = 0; u < GN1_UNROLL; ++u) {
      const int pp = p + u * ppb;
      if (active && pp < rows) {
        gn1_slab[pp * wpp + w] = raw[u];
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[u]));
        s += f.x + f.y;
      }
    }
  }
  const float cnt = static_cast<float>(hw) * static_cast<float>(cpg);
  const float s_cta = gn1_block_sum(s, s_red[0]);
  if (threadIdx.x == 0) s_part[0] = s_cta;
  cluster.sync();
  float tot = 0.f;
  for (int r = 0; r < CL; ++r) tot += *cluster.map_shared_rank(&s_part[0], r);
  const float mean = tot / cnt;

  float q = 0.f;
  if (active) {
    for (int pp = p0; pp < rows; pp += ppb) {
      const uint32_t r = gn1_slab[pp * wpp + w];
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r));
      const float d0 = f.x - mean, d1 = f.y - mean;
      q += d0 * d0 + d1 * d1;
    }
  }
  const float q_cta = gn1_block_sum(q, s_red[1]);
  if (threadIdx.x == 0) s_part[1] = q_cta;
  cluster.sync();
  float qt = 0.f;
  for (int r = 0; r < CL; ++r) qt += *cluster.map_shared_rank(&s_part[1], r);
  const float rstd = rsqrtf(qt / cnt + eps);  // biased variance, as torch GroupNorm

  if (active) {
    const float2 gm = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(gamma + c));
    const float2 bt = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(beta + c));
    const float a0 = gm.x * rstd, a1 = gm.y * rstd;
    const float b0 = bt.x - mean * a0, b1 = bt.y - mean * a1;
    __nv_bfloat16* obase = out + (static_cast<int64_t>(n) * hw + r0) * C + c;
#pragma unroll 4
    for (int pp = p0; pp < rows; pp += ppb) {
      const uint32_t r = gn1_slab[pp * wpp + w];
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r));
      float y0 = f.x * a0 + b0, y1 = f.y * a1 + b1;
      if (silu) {
        y0 = silu_fast(y0);
        y1 = silu_fast(y1);
      }
      *reinterpret_cast<__nv_bfloat162*>(obase + static_cast<int64_t>(pp) * C) = __floats2bfloat162_rn(y0, y1);
    }
  }
  cluster.sync();  // no CTA may exit while a peer can still read its s_part
}

// Window-major cluster GroupNorm for FULL batches (>= 4 images). The group-major kernels above touch 20-80-byte
// segments of every pixel with 4-byte accesses: 2.6x sector over-fetch and ~3x the instructions (ncu at 8x4096x320:
// issue-bound, 27 us for 6.5 us of HBM traffic). Here a cluster of CL CTAs owns a WINDOW of CW = lcm(cpg, 16) channels
// (whole groups, 32-byte-sector aligned) of one image, each CTA a range of pixels: every access is a 16-byte vector
// and whole sectors are used. The window's slab is read from HBM once into shared memory; per-channel sums are reduced
// per group inside the CTA, the per-group partials of the CL CTAs are combined over distributed shared memory (rank
// order -> deterministic), then the exact centred second moment the same way, then normalise (+SiLU) and store.
// Thread (tx, ty) = (tid % vpr, tid / vpr) owns channel octet tx of every TY-th row: fixed channels per thread.
// (One cluster per IMAGE with all channels needs 16-CTA clusters for the 64x64 sites; B200 hosts only 7 of those at
// once — measured with cudaOccupancyMaxActiveClusters — so 8 images would take two waves.)
constexpr int GNR_UNROLL = 8;
constexpr int GNR_MAX_THREADS = 320;
constexpr int GNR_MAX_WGROUPS = 8;  // groups per window (CW / cpg)

__global__ void __launch_bounds__(GNR_MAX_THREADS)
gn_window_cluster_kernel(const __nv_bfloat16* __restrict__ x1, int c1, const __nv_bfloat16* __restrict__ x2, int c2,
                         int hw, int groups, int cw, float eps, int silu, const __nv_bfloat16* __restrict__ gamma,
                         const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  const int C = c1 + c2;
  const int vpr = cw >> 3;  // 16-byte vectors per row of the window
  const int cpg = C / groups;
  const int wgroups = cw / cpg;
  const int TY = blockDim.x / vpr;
  const int tx = threadIdx.x % vpr, ty = threadIdx.x / vpr;
  const int n = blockIdx.y;
  const int win = blockIdx.x / CL;
  const int rows_per = (hw + CL - 1) / CL;
  const int r0 = rank * rows_per;
  const int rows = (r0 + rows_per <= hw) ? rows_per : (hw > r0 ? hw - r0 : 0);
  extern __shared__ uint4 gnr_slab[];                                                        // [rows_per][vpr]
  float* s_acc = reinterpret_cast<float*>(gnr_slab + static_cast<size_t>(rows_per) * vpr);   // [TY][cw]
  __shared__ float s_part[2][GNR_MAX_WGROUPS];  // this CTA's per-group partials (read by the whole cluster)
  __shared__ float s_stat[2][GNR_MAX_WGROUPS];  // mean, rstd
  const int cl0 = tx * 8;            // channel inside the window
  const int c = win * cw + cl0;      // channel inside the tensor
  GnSrc src = gn_src(x1, c1, x2, c2, n, hw, c);
  src.base += static_cast<int64_t>(r0) * src.cs;
  const float cnt = static_cast<float>(hw) * static_cast<float>(cpg);

  // ---- pass 1: HBM -> shared memory, per-channel sums
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int r = ty; r < rows; r += TY * GNR_UNROLL) {
    uint4 v[GNR_UNROLL];
#pragma unroll
    for (int u = 0; u < GNR_UNROLL; ++u) {
      const int rr = r + u * TY;
      v[u] = rr < rows ? *reinterpret_cast<const uint4*>(src.base + static_cast<int64_t>(rr) * src.cs) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < GNR_UNROLL; ++u) {
      const int rr = r + u * TY;
      if (rr < rows) {
        gnr_slab[static_cast<size_t>(rr) * vpr + tx] = v[u];
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k];
      }
    }
  }
  auto reduce_groups = [&](int which) {
    // per-channel partials of the TY row phases -> per-group partial of this CTA (fixed order), published for the cluster
    float4* dst = reinterpret_cast<float4*>(s_acc + static_cast<size_t>(ty) * cw + cl0);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    __syncthreads();
    // warp w reduces group w: lanes stride over the TY x cpg values, fixed-order shuffle tree
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (wid < wgroups) {
      float t = 0.f;
      const int total = TY * cpg;
      for (int i = lane; i < total; i += 32) t += s_acc[static_cast<size_t>(i / cpg) * cw + wid * cpg + i % cpg];
      t = warp_sum(t);
      if (lane == 0) s_part[which][wid] = t;
    }
    cluster.sync();
    if (threadIdx.x < wgroups) {
      float tot = 0.f;
      for (int r = 0; r < CL; ++r) tot += *cluster.map_shared_rank(&s_part[which][threadIdx.x], r);
      s_stat[which][threadIdx.x] = which == 0 ? tot / cnt : rsqrtf(tot / cnt + eps);  // biased variance, as torch
    }
    __syncthreads();
  };
  reduce_groups(0);

  // ---- pass 2: centred second moment from the on-chip copy
  float mean[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mean[k] = s_stat[0][(cl0 + k) / cpg];
    acc[k] = 0.f;
  }
  for (int r = ty; r < rows; r += TY) {
    float f[8];
    unpack8(gnr_slab[static_cast<size_t>(r) * vpr + tx], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float d = f[k] - mean[k];
      acc[k] = fmaf(d, d, acc[k]);
    }
  }
  reduce_groups(1);

  // ---- pass 3: normalise (+SiLU), 16-byte stores of whole sectors
  float a[8], b[8];
  {
    float gm[8], bt[8];
    unpack8(*reinterpret_cast<const uint4*>(gamma + c), gm);
    unpack8(*reinterpret_cast<const uint4*>(beta + c), bt);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a[k] = gm[k] * s_stat[1][(cl0 + k) / cpg];
      b[k] = bt[k] - mean[k] * a[k];
    }
  }
  __nv_bfloat16* obase = out + (static_cast<int64_t>(n) * hw + r0) * C + c;
#pragma unroll 2
  for (int r = ty; r < rows; r += TY) {
    float f[8];
    unpack8(gnr_slab[static_cast<size_t>(r) * vpr + tx], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      f[k] = fmaf(f[k], a[k], b[k]);
      if (silu) f[k] = silu_fast(f[k]);
    }
    *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(r) * C) = pack8(f);
  }
  cluster.sync();  // no CTA may exit while a peer can still read its s_part
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers, two-pass (exact mean, then centered variance). Each warp walks
// rows with a grid stride and loads its next row before it reduces the current one, so a row's memory round trip
// overlaps the previous row's arithmetic (one-row-per-warp blocks spent most of their life in launch + first load).
// ------------------------------------------------------------------------------------------------

template <int VPL>  // 16-byte vectors per lane: C = 32 * 8 * VPL at most
__global__ void __launch_bounds__(256)
layernorm_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ gamma,
                      const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ out, int64_t ldo, int M,
                      int C, float eps) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * 8;
  int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nvec = C / 8;
  const float inv_c = 1.f / static_cast<float>(C);
  uint4 raw[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + i * 32;
    raw[i] = v < nvec ? *reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx + v * 8)
                      : make_uint4(0u, 0u, 0u, 0u);
  }
  while (true) {
    const int next = row + stride;
    uint4 nraw[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      nraw[i] = (next < M && v < nvec) ? *reinterpret_cast<const uint4*>(x + static_cast<int64_t>(next) * ldx + v * 8)
                                       : make_uint4(0u, 0u, 0u, 0u);
    }
    float f[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      unpack8(raw[i], f[i]);  // lanes beyond the row hold zeros: they add nothing to the sum
#pragma unroll
      for (int k = 0; k < 8; ++k) s += f[i][k];
    }
    const float mean = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + i * 32 < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float d = f[i][k] - mean;
          q += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float gm[8], bt[8], o[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + v * 8)), gm);
        unpack8(__ldg(reinterpret_cast<const uint4*>(beta + v * 8)), bt);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (f[i][k] - mean) * rstd * gm[k] + bt[k];
        *reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * ldo + v * 8) = pack8(o);
      }
    }
    if (next >= M) break;
    row = next;
#pragma unroll
    for (int i = 0; i < VPL; ++i) raw[i] = nraw[i];
  }
}

// small fp32 rows (camera encoder): one warp per row, any C
__global__ void layernorm_f32_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                                     const __nv_bfloat16* __restrict__ beta, float* __restrict__ out, int M, int C,
                                     float eps, int silu) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + static_cast<int64_t>(row) * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mean = warp_sum(s) / C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xr[c] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  for (int c = lane; c < C; c += 32) {
    float v = (xr[c] - mean) * rstd * __bfloat162float(gamma[c]) + __bfloat162float(beta[c]);
    if (silu) v = silu_f(v);
    out[static_cast<int64_t>(row) * C + c] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// reference-feature normalisation (src/models/attention.py:95-103)
// ------------------------------------------------------------------------------------------------
// mode "pixel": x is NHWC [B, HW, C]; statistics per pixel over (B, C). One warp per pixel.
__global__ void __launch_bounds__(256)
refnorm_pixel_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int HW, int C) {
  const int pix = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pix >= HW) return;
  const int nvec = C / 8;
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    const __nv_bfloat16* r = x + (static_cast<int64_t>(b) * HW + pix) * C;
    for (int v = lane; v < nvec; v += 32) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(r + v * 8), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += f[k];
    }
  }
  const float cnt = static_cast<float>(B) * C;
  const float mean = warp_sum(s) / cnt;
  float q = 0.f;
  for (int b = 0; b < B; ++b) {
    const __nv_bfloat16* r = x + (static_cast<int64_t>(b) * HW + pix) * C;
    for (int v = lane; v < nvec; v += 32) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(r + v * 8), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = f[k] - mean;
        q += d * d;
      }
    }
  }
  const float stdv = fmaxf(sqrtf(warp_sum(q) / (cnt - 1.f)), 1e-6f);  // unbiased, clamped
  const float scale = 0.5f / stdv;
  for (int b = 0; b < B; ++b) {
    const int64_t off = (static_cast<int64_t>(b) * HW + pix) * C;
    for (int v = lane; v < nvec; v += 32) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(x + off + v * 8), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = (f[k] - mean) * scale;
      *reinterpret_cast<uint4*>(out + off + v * 8) = pack8(f);
    }
  }
}

// mode "channel": x is [B, S, C] (== [B*S, C]); statistics per channel over all B*S rows.
// pass 1: per-chunk column sums -> pass 2: means, per-chunk centered sums of squares -> pass 3: apply.
constexpr int RN_CHUNKS = 64;
__global__ void __launch_bounds__(256)
refnorm_col_sum_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int C, const float* __restrict__ mean_in,
                       float* __restrict__ partial /*[chunks][C]*/) {
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (c >= C) return;
  const int chunk = blockIdx.x, nchunks = gridDim.x;
  const int64_t r0 = rows * chunk / nchunks, r1 = rows * (chunk + 1) / nchunks;
  float m[8], acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    m[k] = mean_in ? mean_in[c + k] : 0.f;
    acc[k] = 0.f;
  }
  for (int64_t r = r0; r < r1; ++r) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(x + r * C + c), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float d = f[k] - m[k];
      acc[k] += mean_in ? d * d : d;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) partial[static_cast<int64_t>(chunk) * C + c + k] = acc[k];
}
// reduce partial[chunks][C] -> out[C] = f(sum): mode 0: mean = sum / rows ; mode 1: 0.5 / clamp(sqrt(rep * sum / (rep * rows - 1)))
// (rep identical copies of the rows: the statistics of the replicated tensor without materialising it)
__global__ void refnorm_col_finalize_kernel(const float* __restrict__ partial, int chunks, int C, double rows, int mode,
                                            float* __restrict__ out, double rep) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0;
  for (int k = 0; k < chunks; ++k) s += partial[static_cast<int64_t>(k) * C + c];
  if (mode == 0) {
    out[c] = static_cast<float>(s / rows);
  } else {
    const float stdv = fmaxf(static_cast<float>(sqrt(rep * s / (rep * rows - 1.0))), 1e-6f);
    out[c] = 0.5f / stdv;
  }
}
__global__ void __launch_bounds__(256)
refnorm_col_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t nvec_total,
                         int C, const float* __restrict__ mean, const float* __restrict__ scale) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nvec_total) return;
  const int c = static_cast<int>((i * 8) % C);
  float f[8];
  unpack8(*reinterpret_cast<const uint4*>(x + i * 8), f);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = (f[k] - mean[c + k]) * scale[c + k];
  *reinterpret_cast<uint4*>(out + i * 8) = pack8(f);
}

}  // namespace mvd

extern "C" {

int64_t mvd_groupnorm_workspace_floats(int n_img, int hw, int groups) {
  // [tickets: n_img uint32, padded to 64] [stats: n_img*groups*2] [partials: n_img*chunks*groups*2]
  (void)hw;
  return 64 + static_cast<int64_t>(n_img > 64 ? n_img : 0) + static_cast<int64_t>(n_img) * groups * 2 +
         static_cast<int64_t>(n_img) * mvd::GN_MAX_CHUNKS * groups * 2;
}

int mvd_groupnorm_bf16(const void* x1, int c1, const void* x2, int c2, const void* gamma, const void* beta, void* out,
                       int n_img, int hw, int groups, float eps, int silu, float* workspace, int64_t workspace_floats,
                       void* stream) {
  using namespace mvd;
  const int C = c1 + ((x2 != nullptr) ? c2 : 0);
  if (x2 == nullptr) c2 = 0;
  MVD_CHECK(n_img > 0 && hw > 0, "groupnorm: empty problem");
  MVD_CHECK(groups > 0 && groups <= GN_MAX_GROUPS && C % groups == 0, "groupnorm: groups=%d C=%d unsupported", groups, C);
  MVD_CHECK(c1 % 8 == 0 && c2 % 8 == 0 && C / 8 <= GN_THREADS * GN_MAX_VPT,
            "groupnorm: channel counts must be multiples of 8 and C <= 4096 (C=%d)", C);
  MVD_CHECK(workspace_floats >= mvd_groupnorm_workspace_floats(n_img, hw, groups), "groupnorm: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto a1 = static_cast<const __nv_bfloat16*>(x1);
  auto a2 = static_cast<const __nv_bfloat16*>(x2);
  auto gm = static_cast<const __nv_bfloat16*>(gamma);
  auto bt = static_cast<const __nv_bfloat16*>(beta);
  auto oo = static_cast<__nv_bfloat16*>(out);
  {
    // Single launch when a group's slab (hw x cpg bf16) fits in shared memory twice per SM; larger slabs stream
    // through the two-kernel path. MVD_GN_TWO_KERNEL=1 forces the latter, MVD_GN_FUSED_MAX_KB lowers the limit.
    static const bool two_kernel_only = [] {
      const char* e = getenv("MVD_GN_TWO_KERNEL");
      return e != nullptr && e[0] == '1';
    }();
    static const int64_t fused_max = [] {
      const char* e = getenv("MVD_GN_FUSED_MAX_KB");
      const int64_t v = e != nullptr ? static_cast<int64_t>(atoi(e)) * 1024 : GN1_MAX_SMEM;
      return v > GN1_MAX_SMEM ? static_cast<int64_t>(GN1_MAX_SMEM) : v;
    }();
    const int cpg = C / groups;
    // full batches: window-major cluster kernel (16-byte accesses to sector-aligned channel windows). MVD_GN_ROWS=0
    // disables it.
    static const bool rows_on = [] {
      const char* e = getenv("MVD_GN_ROWS");
      return e == nullptr || e[0] != '0';
    }();
    // Measured (profiles/r2_gn_window.txt, 8 images): 22.8 -> 15.3 us at 4096 x 320, 18.2 -> 16.4 us at 1024 x 1280, but
    // slower than the group-major kernel for smaller images (three cluster-wide phases of fixed latency): used from
    // 1.3 M elements per image. MVD_GN_ROWS=2 forces it wherever it fits (tests).
    static const bool rows_forced = [] {
      const char* e = getenv("MVD_GN_ROWS");
      return e != nullptr && e[0] == '2';
    }();
    if (rows_on && !two_kernel_only && n_img >= 4 && hw >= 64 &&
        (rows_forced || static_cast<int64_t>(hw) * C >= 4096 * 320)) {
      auto lcm = [](int a, int b) {
        int x = a, y = b;
        while (y) {
          const int t = x % y;
          x = y;
          y = t;
        }
        return a / x * b;
      };
      // window = whole groups, preferably whole 32-byte sectors; cluster = fewest CTAs whose slabs fit twice per SM
      for (int align = 16; align >= 8; align >>= 1) {
        const int cw = lcm(cpg, align);
        if (C % cw != 0 || cw / cpg > GNR_MAX_WGROUPS || cw / 8 > GNR_MAX_THREADS) continue;
        const int vpr = cw / 8;
        const int threads = GNR_MAX_THREADS / vpr * vpr;
        const int windows = C / cw;
        // ALL CTAs of the launch must be resident at once (a second wave costs a whole CTA latency: measured 43-60 us
        // with 512 one-per-SM CTAs): two CTAs per SM (<= 320 threads at 86 registers, slabs <= 104 KB), at most
        // 2 x SMs CTAs, the largest cluster that respects this (>= 32 rows per CTA)
        int cl = 1;
        while (cl < 8 && windows * (2 * cl) * n_img <= 2 * sm_count() && hw / (2 * cl) >= 32) cl <<= 1;
        if (windows * cl * n_img > 2 * sm_count()) continue;
        const size_t smem = static_cast<size_t>((hw + cl - 1) / cl) * cw * 2 + static_cast<size_t>(threads / vpr) * cw * 4;
        if (smem > 104 * 1024) continue;
        MVD_CUDA(cudaFuncSetAttribute(gn_window_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(windows * cl, n_img);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cl;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        int na = 1;
        if (pdl_enabled()) {
          attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          attr[na].val.programmaticStreamSerializationAllowed = 1;
          ++na;
        }
        cfg.attrs = attr;
        cfg.numAttrs = na;
        MVD_CUDA(cudaLaunchKernelEx(&cfg, gn_window_cluster_kernel, a1, c1, a2, c2, hw, groups, cw, eps, silu, gm, bt, oo));
        MVD_CUDA(cudaGetLastError());
        count_launches(1);
        return MVD_OK;
      }
    }
    const int64_t slab = static_cast<int64_t>(hw) * cpg * 2;
    // small batches: CL CTAs (a thread-block cluster) per slab so that the launch still covers the machine
    int cl = 1;
    static const bool cluster_on = [] {
      const char* e = getenv("MVD_GN_CLUSTER");
      return e == nullptr || e[0] != '0';
    }();
    if (cluster_on && hw >= 256) {
      const int want = sm_count() / (groups * n_img);
      cl = want >= 8 ? 8 : want >= 4 ? 4 : want >= 2 ? 2 : 1;
      while (cl > 1 && hw / cl < 64) cl >>= 1;
    }
    const int64_t slab_cta = static_cast<int64_t>((hw + cl - 1) / cl) * cpg * 2;
    if (!two_kernel_only && cpg % 2 == 0 && cpg / 2 <= 128 && slab_cta <= fused_max) {
      // per call: cheap, and correct for every device a process may touch
      MVD_CUDA(cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GN1_MAX_SMEM));
      if (cl > 1) {
        const int rows_per = (hw + cl - 1) / cl;
        const size_t smem = static_cast<size_t>(rows_per) * cpg * 2;
        const int64_t words = static_cast<int64_t>(smem) / 4;
        const int threads1 = words >= 8192 ? GN1_THREADS : (words >= 1024 ? 256 : 128);
        MVD_CUDA(cudaFuncSetAttribute(gn_fused_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GN1_MAX_SMEM));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(groups * cl, n_img);
        cfg.blockDim = dim3(threads1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cl;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        int na = 1;
        if (pdl_enabled()) {
          attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          attr[na].val.programmaticStreamSerializationAllowed = 1;
          ++na;
        }
        cfg.attrs = attr;
        cfg.numAttrs = na;
        MVD_CUDA(cudaLaunchKernelEx(&cfg, gn_fused_cluster_kernel, a1, c1, a2, c2, hw, groups, eps, silu, gm, bt, oo));
        MVD_CUDA(cudaGetLastError());
        count_launches(1);
        return MVD_OK;
      }
      const int64_t words = slab / 4;
      const int threads1 = words >= 8192 ? GN1_THREADS : (words >= 1024 ? 256 : 128);
      MVD_CUDA(launch_pdl(gn_fused_kernel, dim3(groups, n_img), dim3(threads1), static_cast<size_t>(slab), st, a1, c1,
                          a2, c2, hw, groups, eps, silu, gm, bt, oo));
      MVD_CUDA(cudaGetLastError());
      count_launches(1);
      return MVD_OK;
    }
  }
  const int nvec = C / 8;
  const int vpt = nvec > GN_THREADS ? 2 : 1;
  const int TX = (nvec + vpt - 1) / vpt;
  int TY = GN_THREADS / TX;
  if (TY < 1) TY = 1;
  const int chunks = gn_chunks(hw, n_img, TY);
  const int threads = (TX * TY + 31) / 32 * 32;  // whole warps; threads with ty >= TY idle in the row loops
  size_t stats_smem = static_cast<size_t>(2) * TY * C * sizeof(float);
  if (stats_smem < 2 * 8 * 32 * sizeof(double)) stats_smem = 2 * 8 * 32 * sizeof(double);
  MVD_CHECK(stats_smem <= 48 * 1024, "groupnorm: C=%d needs too much shared memory", C);
  // workspace carve-up; the ticket counters must be zero before the FIRST use (the kernel re-arms them itself)
  const int ticket_slots = n_img > 64 ? n_img + 64 : 64;
  unsigned int* tickets = reinterpret_cast<unsigned int*>(workspace);
  float* stats = workspace + ticket_slots;
  float* partial = stats + static_cast<int64_t>(n_img) * groups * 2;
  if (vpt == 1)
    MVD_CUDA(launch_pdl(gn_stats_kernel<1>, dim3(chunks, n_img), dim3(threads), stats_smem, st, a1, c1, a2, c2, hw,
                        groups, TX, TY, eps, partial, stats, tickets));
  else
    MVD_CUDA(launch_pdl(gn_stats_kernel<2>, dim3(chunks, n_img), dim3(threads), stats_smem, st, a1, c1, a2, c2, hw,
                        groups, TX, TY, eps, partial, stats, tickets));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  if (vpt == 1)
    MVD_CUDA(launch_pdl(gn_apply_kernel<1>, dim3(chunks, n_img), dim3(threads), 0, st, a1, c1, a2, c2, hw, groups, silu,
                        gm, bt, stats, TX, TY, oo));
  else
    MVD_CUDA(launch_pdl(gn_apply_kernel<2>, dim3(chunks, n_img), dim3(threads), 0, st, a1, c1, a2, c2, hw, groups, silu,
                        gm, bt, stats, TX, TY, oo));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_layernorm_bf16(const void* x, int64_t ldx, const void* gamma, const void* beta, void* out, int64_t ldo, int M,
                       int C, float eps, void* stream) {
  using namespace mvd;
  MVD_CHECK(M > 0 && C > 0 && C % 8 == 0 && C <= 2048 && ldx % 8 == 0 && ldo % 8 == 0,
            "layernorm: C (=%d) must be a multiple of 8 and <= 2048, strides multiples of 8", C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vpl = (C / 8 + 31) / 32;
  int blocks = (M + 7) / 8;
  const int resident = sm_count() * (vpl <= 2 ? 4 : 2);  // 60 / 107-128 registers per thread
  if (blocks > resident) blocks = resident;
  auto xx = static_cast<const __nv_bfloat16*>(x);
  auto g = static_cast<const __nv_bfloat16*>(gamma);
  auto b = static_cast<const __nv_bfloat16*>(beta);
  auto o = static_cast<__nv_bfloat16*>(out);
  if (vpl <= 2)
    MVD_CUDA(launch_pdl(layernorm_bf16_kernel<2>, dim3(blocks), dim3(256), 0, st, xx, ldx, g, b, o, ldo, M, C, eps));
  else if (vpl <= 5)
    MVD_CUDA(launch_pdl(layernorm_bf16_kernel<5>, dim3(blocks), dim3(256), 0, st, xx, ldx, g, b, o, ldo, M, C, eps));
  else
    MVD_CUDA(launch_pdl(layernorm_bf16_kernel<8>, dim3(blocks), dim3(256), 0, st, xx, ldx, g, b, o, ldo, M, C, eps));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_layernorm_f32(const float* x, const void* gamma, const void* beta, float* out, int M, int C, float eps,
                      int silu, void* stream) {
  using namespace mvd;
  MVD_CHECK(M > 0 && C > 0, "layernorm_f32: empty problem");
  layernorm_f32_kernel<<<(M + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<const __nv_bfloat16*>(gamma), static_cast<const __nv_bfloat16*>(beta), out, M, C, eps, silu);
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int64_t mvd_refnorm_workspace_floats(int channels) { return static_cast<int64_t>(mvd::RN_CHUNKS + 2) * channels; }

int mvd_refnorm_bf16(const void* x, void* out, int batch, int seq, int channels, int per_pixel, float* workspace,
                     int64_t workspace_floats, void* stream) {
  return mvd_refnorm_replicated_bf16(x, out, batch, seq, channels, per_pixel, 1, workspace, workspace_floats, stream);
}

int mvd_refnorm_replicated_bf16(const void* x, void* out, int batch, int seq, int channels, int per_pixel,
                                int replication, float* workspace, int64_t workspace_floats, void* stream) {
  using namespace mvd;
  MVD_CHECK(replication >= 1 && (replication == 1 || !per_pixel),
            "refnorm: replication applies to the 3-D (per-channel) form only");
  MVD_CHECK(batch > 0 && seq > 0 && channels > 0 && channels % 8 == 0, "refnorm: bad shape B=%d S=%d C=%d", batch, seq,
            channels);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto xx = static_cast<const __nv_bfloat16*>(x);
  auto oo = static_cast<__nv_bfloat16*>(out);
  if (per_pixel) {
    MVD_CHECK(static_cast<int64_t>(batch) * channels > 1, "refnorm: unbiased std needs more than one element");
    refnorm_pixel_kernel<<<(seq + 7) / 8, 256, 0, st>>>(xx, oo, batch, seq, channels);
    MVD_CUDA(cudaGetLastError());
  count_launches(1);
    return MVD_OK;
  }
  MVD_CHECK(workspace_floats >= mvd_refnorm_workspace_floats(channels), "refnorm: workspace too small");
  const int64_t rows = static_cast<int64_t>(batch) * seq;
  MVD_CHECK(rows * replication > 1, "refnorm: unbiased std needs more than one row");
  const int chunks = rows < RN_CHUNKS ? static_cast<int>(rows) : RN_CHUNKS;
  float* partial = workspace;
  float* mean = workspace + static_cast<int64_t>(RN_CHUNKS) * channels;
  float* scale = mean + channels;
  const int tx = 64;
  dim3 grid(chunks, (channels / 8 + tx - 1) / tx);
  refnorm_col_sum_kernel<<<grid, tx, 0, st>>>(xx, rows, channels, nullptr, partial);
  refnorm_col_finalize_kernel<<<(channels + 127) / 128, 128, 0, st>>>(partial, chunks, channels,
                                                                      static_cast<double>(rows), 0, mean, 1.0);
  refnorm_col_sum_kernel<<<grid, tx, 0, st>>>(xx, rows, channels, mean, partial);
  refnorm_col_finalize_kernel<<<(channels + 127) / 128, 128, 0, st>>>(partial, chunks, channels,
                                                                      static_cast<double>(rows), 1, scale,
                                                                      static_cast<double>(replication));
  const int64_t nvec = rows * channels / 8;
  refnorm_col_apply_kernel<<<static_cast<unsigned>((nvec + 255) / 256), 256, 0, st>>>(xx, oo, nvec, channels, mean,
                                                                                      scale);
  MVD_CUDA(cudaGetLastError());
  count_launches(5);
  return MVD_OK;
}

}  // extern "C"
