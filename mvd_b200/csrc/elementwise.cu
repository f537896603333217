// elementwise.cu — bandwidth/latency-bound kernels of the denoise step: camera FiLM, embeddings, skinny
// linears, the 4-channel edge convolutions, layout/resampling kernels and the fused CFG + DDPM step.
// All global accesses on the large tensors are 128-bit and coalesced; reductions use warp shuffles.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "host_common.h"
#include "../../include/mvd_b200.h"

namespace mvd {

__device__ __forceinline__ float ew_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void ew_unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 ew_pack8(const float* f) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
__device__ __forceinline__ float ew_silu(float x) { return x / (1.f + __expf(-x)); }

// ------------------------------------------------------------------------------------------------
// Camera FiLM (src/models/camera_encoder.py:221-234): y = x * (2*sigmoid(s)*strength) + shift*strength
// mod: fp32 [V, 2C] (modulator MLP output: first C = scale logits, last C = shift); sample n uses
// row n % V (CFG halves share the cameras). x/out: NHWC bf16 [N, HW, C].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
film_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, const float* __restrict__ mod,
            int V, int hw, int C, float strength) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float s_coef[];  // [2][C]
  const int n = blockIdx.y;
  const float* m = mod + static_cast<int64_t>(n % V) * 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_coef[c] = 2.f * strength / (1.f + __expf(-m[c]));
    s_coef[C + c] = m[C + c] * strength;
  }
  __syncthreads();
  const int nvec_row = C / 8;
  const int total = hw * nvec_row;  // 16-byte vectors per image (host checks that this fits in 31 bits)
  const int per_block = (total + gridDim.x - 1) / gridDim.x;
  const int i0 = per_block * blockIdx.x;
  const int i1 = (i0 + per_block < total) ? i0 + per_block : total;
  const __nv_bfloat16* xb = x + static_cast<int64_t>(n) * hw * C;
  __nv_bfloat16* ob = out + static_cast<int64_t>(n) * hw * C;
  constexpr int U = 4;  // independent 16-byte loads in flight per thread
  for (int i = i0 + threadIdx.x; i < i1; i += U * blockDim.x) {
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ii = i + u * blockDim.x;
      raw[u] = ii < i1 ? *reinterpret_cast<const uint4*>(xb + static_cast<int64_t>(ii) * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ii = i + u * blockDim.x;
      if (ii < i1) {
        const int c = (ii % nvec_row) * 8;
        float f[8];
        ew_unpack8(raw[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = f[k] * s_coef[c + k] + s_coef[C + c + k];
        *reinterpret_cast<uint4*>(ob + static_cast<int64_t>(ii) * 8) = ew_pack8(f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Skinny linear: out[M, N] = act_out(act_in(x[M, K]) @ w[N, K]^T + b), M <= 16, fp32 activations, bf16 weights.
// One warp per output feature; weight rows are streamed once with 128-bit loads (bandwidth-bound on w).
// Used for TimestepEmbedding, all 22 ResnetBlock2D.time_emb_proj at once, and the CameraEncoder MLPs.
// ------------------------------------------------------------------------------------------------
constexpr int SL_MAX_M = 16;
constexpr int SL_KTILE = 1024;  // K elements of x staged in shared memory at a time (M x 1024 fp32 = 64 KB max)
constexpr int SL_RPW = 4;       // output features per warp (32 per block): amortises the activation staging
template <int MT>               // MT = compile-time bound on M (register accumulators)
__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ w,
                    const __nv_bfloat16* __restrict__ b, float* __restrict__ out, int64_t ldo, int M, int N, int K,
                    int silu_in, int silu_out) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float s_x[];  // [M][kt] activations (SiLU already applied), shared by the 8 warps of the block
  const int n0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * SL_RPW;
  const int lane = threadIdx.x & 31;
  float acc[SL_RPW][MT];
#pragma unroll
  for (int r = 0; r < SL_RPW; ++r)
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[r][m] = 0.f;
  const bool vec = (K & 7) == 0;
  for (int k0 = 0; k0 < K; k0 += SL_KTILE) {
    const int kt = (K - k0 < SL_KTILE) ? K - k0 : SL_KTILE;
    __syncthreads();
    for (int i = threadIdx.x; i < M * kt; i += blockDim.x) {
      const int m = i / kt, k = i % kt;
      float v = x[m * ldx + k0 + k];
      if (silu_in) v = ew_silu(v);
      s_x[m * kt + k] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SL_RPW; ++r) {
      const int n = n0 + r;
      if (n >= N) break;
      const __nv_bfloat16* wr = w + static_cast<int64_t>(n) * K + k0;
      if (vec) {
        for (int k = lane * 8; k < kt; k += 256) {
          float wf[8];
          ew_unpack8(*reinterpret_cast<const uint4*>(wr + k), wf);
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            if (m < M) {
              const float4 x0 = *reinterpret_cast<const float4*>(s_x + m * kt + k);
              const float4 x1 = *reinterpret_cast<const float4*>(s_x + m * kt + k + 4);
              acc[r][m] += x0.x * wf[0] + x0.y * wf[1] + x0.z * wf[2] + x0.w * wf[3] + x1.x * wf[4] + x1.y * wf[5] +
                           x1.z * wf[6] + x1.w * wf[7];
            }
          }
        }
      } else {
        for (int k = lane; k < kt; k += 32) {
          const float wv = __bfloat162float(wr[k]);
#pragma unroll
          for (int m = 0; m < MT; ++m)
            if (m < M) acc[r][m] += s_x[m * kt + k] * wv;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < SL_RPW; ++r) {
    const int n = n0 + r;
    if (n >= N) break;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (m < M) {
        float v = ew_warp_sum(acc[r][m]);
        if (lane == 0) {
          if (b != nullptr) v += __bfloat162float(b[n]);
          if (silu_out) v = ew_silu(v);
          out[m * ldo + n] = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Timestep sinusoid (diffusers Timesteps(320, flip_sin_to_cos=True, freq_shift=0)): out[b] = [cos | sin]
// ------------------------------------------------------------------------------------------------
__global__ void timestep_embed_kernel(const float* __restrict__ t, int n_t, float* __restrict__ out, int B, int dim) {
  pdl_wait();
  pdl_launch_dependents();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, j = i % half;
  const float tv = t[n_t == 1 ? 0 : b];
  const float freq = expf(-logf(10000.f) * static_cast<float>(j) / static_cast<float>(half));
  const float ang = tv * freq;
  out[static_cast<int64_t>(b) * dim + j] = cosf(ang);
  out[static_cast<int64_t>(b) * dim + half + j] = sinf(ang);
}

// ------------------------------------------------------------------------------------------------
// Camera front end (src/models/camera_encoder.py:107-120,137-151): relative pose + sinusoidal encoding of T.
//   R_rel = R_t R_s^T (row-major, flattened to 9), T_rel = T_t - R_rel T_s,
//   enc[v, a, :] = [sin(T_a * f_0..P-1) | cos(T_a * f_0..P-1)],  f = exp(linspace(0, ln(max_freq), P))
// src/tgt: fp32 [V, 3, 4]; r_flat: fp32 [V, 9]; enc: fp32 [V, 3 * 2P].
// ------------------------------------------------------------------------------------------------
__global__ void camera_front_kernel(const float* __restrict__ src, const float* __restrict__ tgt,
                                    float* __restrict__ r_flat, float* __restrict__ enc, float* __restrict__ t_rel, int V, int P,
                                    float max_freq) {
  const int v = blockIdx.x;
  __shared__ float R[9], T[3];
  const float* s = src + v * 12;
  const float* t = tgt + v * 12;
  if (threadIdx.x < 9) {
    const int i = threadIdx.x / 3, j = threadIdx.x % 3;
    float acc = 0.f;
    for (int k = 0; k < 3; ++k) acc += t[i * 4 + k] * s[j * 4 + k];  // (R_t R_s^T)_{ij}
    R[threadIdx.x] = acc;
    r_flat[v * 9 + threadIdx.x] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int i = threadIdx.x;
    float acc = 0.f;
    for (int k = 0; k < 3; ++k) acc += R[i * 3 + k] * s[k * 4 + 3];
    T[i] = t[i * 4 + 3] - acc;
    if (t_rel != nullptr) t_rel[v * 3 + i] = T[i];
  }
  __syncthreads();
  const float lmax = logf(max_freq);
  for (int i = threadIdx.x; i < 3 * P; i += blockDim.x) {
    const int a = i / P, j = i % P;
    const float step = (P > 1) ? lmax / static_cast<float>(P - 1) : 0.f;
    const float f = expf(step * static_cast<float>(j));
    const float ang = T[a] * f;
    enc[static_cast<int64_t>(v) * 6 * P + a * 2 * P + j] = sinf(ang);
    enc[static_cast<int64_t>(v) * 6 * P + a * 2 * P + P + j] = cosf(ang);
  }
}

// ------------------------------------------------------------------------------------------------
// conv_in: Conv2d(4, Cout, 3, padding 1) on fp32 NCHW latents, with the camera FiLM on the INPUT latents
// fused in (src/models/mvd_unet.py:256-258 applies the "output" modulator to the input sample) and the CFG
// duplication folded into the index (sample n reads latent n % n_lat). Output NHWC bf16.
// ------------------------------------------------------------------------------------------------
// Lanes are PIXELS (32 consecutive pixels of an image per warp), a thread keeps its pixel's 36 modulated inputs in
// registers and produces Cout/4 output channels, 8 at a time, from an fp32 [36][Cout] copy of the weights in shared
// memory (two broadcast LDS.128 per 8 FMAs). Round 1's form (channel vectors across lanes, 2-way bank conflicts on
// every weight load, [Cout][36] -> [36][Cout] transpose with a 32-way conflict) took 126 us at 8 x 64 x 64.
constexpr int CIN_TP = 64;  // pixels per block (two warps of pixels x four channel quarters)
__global__ void __launch_bounds__(256)
conv_in_kernel(const float* __restrict__ lat, int n_lat, const float* __restrict__ mod /*[V, 8] or null*/, int V,
               float strength, const __nv_bfloat16* __restrict__ w /*[Cout, 3,3, 4]*/,
               const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W, int Cout) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float s_w[];                                             // [36][Cout] fp32, tap-major
  float* s_b = s_w + Cout * 36;                                              // [Cout]
  __nv_bfloat16* s_raw = reinterpret_cast<__nv_bfloat16*>(s_b + Cout);       // [Cout][36] as stored
  for (int i = threadIdx.x; i < Cout * 36 / 8; i += blockDim.x)
    reinterpret_cast<uint4*>(s_raw)[i] = reinterpret_cast<const uint4*>(w)[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) s_b[i] = __bfloat162float(bias[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < Cout * 36; i += blockDim.x) {
    const int tap = i / Cout, co = i - tap * Cout;
    s_w[i] = __bfloat162float(s_raw[co * 36 + tap]);
  }
  const int n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hw = H * W;
  const int pix = blockIdx.x * CIN_TP + (warp >> 2) * 32 + lane;
  const bool valid = pix < hw;
  const int y = valid ? pix / W : 0, x = valid ? pix - (pix / W) * W : 0;
  float sc[4], sh[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    sc[c] = 1.f;
    sh[c] = 0.f;
    if (mod != nullptr) {
      const float* m = mod + (n % V) * 8;
      sc[c] = 2.f * strength / (1.f + __expf(-m[c]));
      sh[c] = m[4 + c] * strength;
    }
  }
  // the pixel's 3x3x4 neighbourhood (zero padding applies to the MODULATED sample), tap-major like the weights
  const float* lb = lat + static_cast<int64_t>(n % n_lat) * 4 * hw;
  float in[36];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xx = x + kx - 1;
      const bool inb = valid && yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        in[(ky * 3 + kx) * 4 + c] = inb ? fmaf(__ldg(lb + static_cast<int64_t>(c) * hw + yy * W + xx), sc[c], sh[c]) : 0.f;
    }
  __syncthreads();
  const int cq = Cout >> 2;  // channels per warp quarter (host checks Cout % 32 == 0)
  const int c_begin = (warp & 3) * cq;
  __nv_bfloat16* op = out + (static_cast<int64_t>(n) * hw + pix) * Cout;
  for (int cv = c_begin; cv < c_begin + cq; cv += 8) {
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = s_b[cv + k];
#pragma unroll
    for (int k = 0; k < 36; ++k) {
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + k * Cout + cv);
      const float4 w1 = *reinterpret_cast<const float4*>(s_w + k * Cout + cv + 4);
      const float v = in[k];
      acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]);
      acc[3] = fmaf(v, w0.w, acc[3]); acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
      acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
    }
    if (valid) *reinterpret_cast<uint4*>(op + cv) = ew_pack8(acc);
  }
}

// ------------------------------------------------------------------------------------------------
// conv_out: Conv2d(Cin, 4, 3, padding 1) on NHWC bf16 (already GroupNorm+SiLU'd) -> fp32 NCHW [N,4,H,W].
// One warp per output pixel; lanes split the channels, shuffle-reduce the 4 outputs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_out_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w /*[4][3][3][Cin]*/,
                const __nv_bfloat16* __restrict__ bias, float* __restrict__ out, int N, int H, int W, int Cin) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __nv_bfloat16 s_wo[];  // [4*9*Cin]
  for (int i = threadIdx.x; i < 36 * Cin / 8; i += blockDim.x)
    reinterpret_cast<uint4*>(s_wo)[i] = reinterpret_cast<const uint4*>(w)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t pix = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (pix >= static_cast<int64_t>(N) * H * W) return;
  const int xq = static_cast<int>(pix % W), yq = static_cast<int>((pix / W) % H), n = static_cast<int>(pix / (W * H));
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int nvec = Cin / 8;
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = yq + tap / 3 - 1, xx = xq + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const __nv_bfloat16* xr = x + ((static_cast<int64_t>(n) * H + yy) * W + xx) * Cin;
    for (int v = lane; v < nvec; v += 32) {
      float f[8];
      ew_unpack8(*reinterpret_cast<const uint4*>(xr + v * 8), f);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        float wf[8];
        ew_unpack8(*reinterpret_cast<const uint4*>(s_wo + (o * 9 + tap) * Cin + v * 8), wf);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[o] += f[k] * wf[k];
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const float v = ew_warp_sum(acc[o]);
    if (lane == 0) out[((static_cast<int64_t>(n) * 4 + o) * H + yq) * W + xq] = v + __bfloat162float(bias[o]);
  }
}

// ------------------------------------------------------------------------------------------------
// nearest x2 upsample, NHWC bf16 (diffusers Upsample2D before its conv)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int N, int H, int W, int cvec) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(N) * 2 * H * 2 * W * cvec;
  if (i >= total) return;
  const int c = static_cast<int>(i % cvec);
  int64_t p = i / cvec;
  const int xo = static_cast<int>(p % (2 * W));
  p /= 2 * W;
  const int yo = static_cast<int>(p % (2 * H));
  const int n = static_cast<int>(p / (2 * H));
  out[i] = x[((static_cast<int64_t>(n) * H + (yo >> 1)) * W + (xo >> 1)) * cvec + c];
}

// First 4 channels of an NHWC bf16 tensor (row stride ld elements) -> fp32 NCHW [N,4,HW]: the tail of conv_out when it
// runs as a 32-column tcgen05 conv (zero-padded weight rows). Thread per pixel: one 8-byte load, four coalesced stores.
__global__ void __launch_bounds__(256)
head4_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, float* __restrict__ out, int N, int64_t HW) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(N) * HW) return;
  const int64_t n = i / HW, px = i - n * HW;
  const uint2 v = *reinterpret_cast<const uint2*>(x + i * ld);
  float* o = out + n * 4 * HW + px;
  o[0] = __uint_as_float(v.x << 16);
  o[HW] = __uint_as_float(v.x & 0xffff0000u);
  o[2 * HW] = __uint_as_float(v.y << 16);
  o[3 * HW] = __uint_as_float(v.y & 0xffff0000u);
}

__global__ void __launch_bounds__(256)
add_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, int64_t nvec) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  float fa[8], fb[8];
  ew_unpack8(a[i], fa);
  ew_unpack8(b[i], fb);
#pragma unroll
  for (int k = 0; k < 8; ++k) fa[k] += fb[k];
  out[i] = ew_pack8(fa);
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    *reinterpret_cast<__nv_bfloat162*>(out + i) = __floats2bfloat162_rn(x[i], x[i + 1]);
  } else if (i < n) {
    out[i] = __float2bfloat16(x[i]);
  }
}

// [B, C, S] (NCHW, any float type) <-> [B, S, C] bf16 through a 32x32 smem tile
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
transpose_bcs_kernel(const TIn* __restrict__ x, TOut* __restrict__ out, int R, int Cc) {
  // x: [B][R][Cc] -> out: [B][Cc][R]
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const TIn* xb = x + static_cast<int64_t>(b) * R * Cc;
  TOut* ob = out + static_cast<int64_t>(b) * R * Cc;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[i][threadIdx.x] = static_cast<float>(xb[static_cast<int64_t>(r) * Cc + c]);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) ob[static_cast<int64_t>(c) * R + r] = static_cast<TOut>(tile[threadIdx.x][i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Fused classifier-free guidance + DDPM (v-prediction, fixed_small variance) step
// (src/models/pipeline.py:156-158,161; diffusers DDPMScheduler.step).
//   v = v_u + g (v_c - v_u);  x0 = sqrt(abar_t) x - sqrt(1-abar_t) v
//   x_prev = c_x0 * x0 + c_xt * x + sigma * noise
// model_out: fp32 [cfg * n, ...] (uncond first, cond second when cfg == 2); latents fp32 updated in place.
// ------------------------------------------------------------------------------------------------
// One arithmetic definition for both step kernels (no FMA contraction: bit-identical to each other, and the same
// operation order as the scheduler's unfused fp32 tensor expressions).
__device__ __forceinline__ float ddpm_update(float v_u, float v_c, int cfg, float guidance, float x, float sa, float sb,
                                             float c0, float ct) {
  float v = v_u;
  if (cfg == 2) v = __fadd_rn(v_u, __fmul_rn(guidance, __fsub_rn(v_c, v_u)));
  const float x0 = __fsub_rn(__fmul_rn(sa, x), __fmul_rn(sb, v));
  return __fadd_rn(__fmul_rn(c0, x0), __fmul_rn(ct, x));
}

__global__ void __launch_bounds__(256)
cfg_ddpm_step_kernel(const float* __restrict__ model_out, float* __restrict__ latents, const float* __restrict__ noise,
                     int64_t n, int cfg, float guidance, float sqrt_abar, float sqrt_1m_abar, float c_x0, float c_xt,
                     float sigma) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float xp = ddpm_update(model_out[i], cfg == 2 ? model_out[n + i] : 0.f, cfg, guidance, latents[i], sqrt_abar,
                         sqrt_1m_abar, c_x0, c_xt);
  if (noise != nullptr && sigma != 0.f) xp = __fadd_rn(xp, __fmul_rn(sigma, noise[i]));
  latents[i] = xp;
}

// Device-table variant for CUDA-graph replay of the whole sampling loop: every per-step scalar is read from
// coef_table[*step_idx] = {t, sqrt_abar, sqrt_1m_abar, c_x0, c_xt, sigma, 0, 0}; noise_table is [steps][n].
__global__ void __launch_bounds__(256)
cfg_ddpm_step_table_kernel(const float* __restrict__ model_out, float* __restrict__ latents,
                           const float* __restrict__ noise_table, int64_t n, int cfg, float guidance,
                           const float* __restrict__ coef_table, const int* __restrict__ step_idx) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = *step_idx;
  const float* c = coef_table + static_cast<int64_t>(s) * 8;
  float xp = ddpm_update(model_out[i], cfg == 2 ? model_out[n + i] : 0.f, cfg, guidance, latents[i], c[1], c[2], c[3],
                         c[4]);
  if (noise_table != nullptr && c[5] != 0.f) xp = __fadd_rn(xp, __fmul_rn(c[5], noise_table[static_cast<int64_t>(s) * n + i]));
  latents[i] = xp;
}
__global__ void advance_step_kernel(int* step_idx, const float* coef_table, float* timestep_out, int n_steps) {
  pdl_wait();
  pdl_launch_dependents();
  int s = *step_idx + 1;
  if (s >= n_steps) s = 0;  // wrap: the loop can be replayed
  *step_idx = s;
  timestep_out[0] = coef_table[static_cast<int64_t>(s) * 8];
}

// Same, plus: copies row `s` (the NEW step) of row_table [n_steps][row_len] into row_out — the per-schedule table of
// everything that depends on the timestep only (time-embedding MLP + the 22 time_emb_proj outputs), so that the
// per-step graph carries no timestep arithmetic at all. One block; every thread reads the old counter before thread
// 0 overwrites it.
__global__ void __launch_bounds__(1024) advance_step_rows_kernel(int* step_idx, const float* coef_table,
                                                                 float* timestep_out, int n_steps,
                                                                 const float4* row_table, float4* row_out, int row_len4) {
  pdl_wait();
  pdl_launch_dependents();
  int s = *step_idx + 1;
  if (s >= n_steps) s = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    *step_idx = s;
    timestep_out[0] = coef_table[static_cast<int64_t>(s) * 8];
  }
  const float4* src = row_table + static_cast<int64_t>(s) * row_len4;
  for (int i = threadIdx.x; i < row_len4; i += blockDim.x) row_out[i] = src[i];
}

}  // namespace mvd

extern "C" {

int mvd_film_bf16(const void* x, void* out, const float* mod, int n_img, int n_cam, int hw, int channels,
                  float strength, void* stream) {
  using namespace mvd;
  MVD_CHECK(n_img > 0 && n_cam > 0 && hw > 0 && channels % 8 == 0 && channels <= 4096, "film: bad shape C=%d",
            channels);
  MVD_CHECK(static_cast<int64_t>(hw) * channels / 8 < (1ll << 30), "film: image too large (hw=%d C=%d)", hw, channels);
  // ~4 batches of 4 vectors per thread; at most ~8 resident blocks per SM over the whole grid
  int bx = static_cast<int>((static_cast<int64_t>(hw) * channels / 8 + 4095) / 4096);
  if (bx < 1) bx = 1;
  while (bx > 1 && static_cast<int64_t>(bx) * n_img > 1184) bx = (bx + 1) / 2;
  MVD_CUDA(launch_pdl(film_kernel, dim3(bx, n_img), dim3(256), 2 * channels * sizeof(float),
                      static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x),
                      static_cast<__nv_bfloat16*>(out), mod, n_cam, hw, channels, strength));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_small_linear_f32(const float* x, int64_t ldx, const void* w, const void* bias, float* out, int64_t ldo, int M,
                         int N, int K, int silu_in, int silu_out, void* stream) {
  using namespace mvd;
  MVD_CHECK(M > 0 && M <= SL_MAX_M && N > 0 && K > 0, "small_linear: M must be in [1,16] (M=%d N=%d K=%d)", M, N, K);
  MVD_CHECK((K & 7) != 0 || ((reinterpret_cast<uintptr_t>(w) & 15) == 0), "small_linear: w must be 16-byte aligned");
  const int kt = K < SL_KTILE ? K : SL_KTILE;
  const size_t smem = static_cast<size_t>(M) * kt * sizeof(float);
  // per call: cheap, and correct for every device a process may touch
  MVD_CUDA(cudaFuncSetAttribute(small_linear_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                SL_MAX_M * SL_KTILE * static_cast<int>(sizeof(float))));
  MVD_CUDA(cudaFuncSetAttribute(small_linear_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                SL_MAX_M * SL_KTILE * static_cast<int>(sizeof(float))));
  const int blocks = (N + 8 * SL_RPW - 1) / (8 * SL_RPW);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto ww = static_cast<const __nv_bfloat16*>(w);
  auto bb = static_cast<const __nv_bfloat16*>(bias);
  if (M <= 8)
    MVD_CUDA(launch_pdl(small_linear_kernel<8>, dim3(blocks), dim3(256), smem, st, x, ldx, ww, bb, out, ldo, M, N, K,
                        silu_in, silu_out));
  else
    MVD_CUDA(launch_pdl(small_linear_kernel<16>, dim3(blocks), dim3(256), smem, st, x, ldx, ww, bb, out, ldo, M, N, K,
                        silu_in, silu_out));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_timestep_embedding_f32(const float* timesteps, int n_timesteps, float* out, int batch, int dim, void* stream) {
  using namespace mvd;
  MVD_CHECK(batch > 0 && dim > 0 && dim % 2 == 0 && (n_timesteps == 1 || n_timesteps == batch),
            "timestep_embedding: bad shape");
  const int total = batch * dim / 2;
  MVD_CUDA(launch_pdl(timestep_embed_kernel, dim3((total + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream),
                      timesteps, n_timesteps, out, batch, dim));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_camera_front_f32(const float* source_cam, const float* target_cam, float* r_flat, float* t_enc, float* t_rel,
                         int n_views, int pos_enc_dim, float max_freq, void* stream) {
  using namespace mvd;
  MVD_CHECK(n_views > 0 && pos_enc_dim > 0, "camera_front: bad shape");
  camera_front_kernel<<<n_views, 128, 0, static_cast<cudaStream_t>(stream)>>>(source_cam, target_cam, r_flat, t_enc,
                                                                              t_rel, n_views, pos_enc_dim, max_freq);
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_conv_in_f32_bf16(const float* latents, int n_latents, const float* mod, int n_cam, float strength,
                         const void* w, const void* bias, void* out, int n_img, int h, int wdt, int c_out,
                         void* stream) {
  using namespace mvd;
  MVD_CHECK(n_img > 0 && n_latents > 0 && h > 0 && wdt > 0 && c_out % 8 == 0 && c_out <= 1024,
            "conv_in: bad shape Cout=%d", c_out);
  MVD_CHECK(c_out % 32 == 0, "conv_in: Cout (=%d) must be a multiple of 32", c_out);
  const size_t smem = static_cast<size_t>(c_out) * 37 * sizeof(float) + static_cast<size_t>(c_out) * 36 * 2;
  // per call: cheap, and correct for every device a process may touch
  MVD_CUDA(cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  MVD_CUDA(launch_pdl(conv_in_kernel, dim3((h * wdt + CIN_TP - 1) / CIN_TP, n_img), dim3(256), smem,
                      static_cast<cudaStream_t>(stream), latents, n_latents, mod, n_cam > 0 ? n_cam : 1, strength,
                      static_cast<const __nv_bfloat16*>(w), static_cast<const __nv_bfloat16*>(bias),
                      static_cast<__nv_bfloat16*>(out), h, wdt, c_out));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_conv_out_bf16_f32(const void* x, const void* w, const void* bias, float* out, int n_img, int h, int wdt,
                          int c_in, void* stream) {
  using namespace mvd;
  MVD_CHECK(n_img > 0 && h > 0 && wdt > 0 && c_in % 8 == 0 && c_in <= 1280, "conv_out: bad shape Cin=%d", c_in);
  const size_t smem = static_cast<size_t>(36) * c_in * sizeof(__nv_bfloat16);
  // per call: cheap, and correct for every device a process may touch
  MVD_CUDA(cudaFuncSetAttribute(conv_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  const int64_t pix = static_cast<int64_t>(n_img) * h * wdt;
  MVD_CUDA(launch_pdl(conv_out_kernel, dim3(static_cast<unsigned>((pix + 7) / 8)), dim3(256), smem,
                      static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x),
                      static_cast<const __nv_bfloat16*>(w), static_cast<const __nv_bfloat16*>(bias), out, n_img, h, wdt,
                      c_in));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_upsample_nearest2x_bf16(const void* x, void* out, int n_img, int h, int wdt, int channels, void* stream) {
  using namespace mvd;
  MVD_CHECK(n_img > 0 && h > 0 && wdt > 0 && channels % 8 == 0, "upsample: bad shape");
  const int64_t total = static_cast<int64_t>(n_img) * 4 * h * wdt * (channels / 8);
  MVD_CUDA(launch_pdl(upsample2x_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0,
                      static_cast<cudaStream_t>(stream), static_cast<const uint4*>(x), static_cast<uint4*>(out), n_img, h,
                      wdt, channels / 8));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_head4_to_nchw_f32(const void* x, int64_t ld, float* out, int n_img, int64_t hw, void* stream) {
  using namespace mvd;
  MVD_CHECK(n_img > 0 && hw > 0 && ld >= 4 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0,
            "head4_to_nchw: rows must be 8-byte aligned with a stride multiple of 4 elements");
  const int64_t total = static_cast<int64_t>(n_img) * hw;
  MVD_CUDA(launch_pdl(head4_to_nchw_f32_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0,
                      static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), ld, out, n_img, hw));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream) {
  using namespace mvd;
  MVD_CHECK(n > 0 && n % 8 == 0, "add: element count must be a positive multiple of 8");
  add_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(out), n / 8);
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_cast_f32_bf16(const float* x, void* out, int64_t n, void* stream) {
  using namespace mvd;
  MVD_CHECK(n > 0, "cast: empty tensor");
  cast_f32_bf16_kernel<<<static_cast<unsigned>(((n + 1) / 2 + 255) / 256), 256, 0,
                         static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(out), n);
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

/* x: [batch, rows, cols] -> out: [batch, cols, rows]; src_dtype/dst_dtype: 0 = fp32, 1 = bf16 */
int mvd_transpose_batched(const void* x, void* out, int batch, int rows, int cols, int src_dtype, int dst_dtype,
                          void* stream) {
  using namespace mvd;
  MVD_CHECK(batch > 0 && rows > 0 && cols > 0 && batch <= 65535, "transpose: bad shape");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch), block(32, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src_dtype == 0 && dst_dtype == 1)
    transpose_bcs_kernel<float, __nv_bfloat16><<<grid, block, 0, st>>>(static_cast<const float*>(x),
                                                                       static_cast<__nv_bfloat16*>(out), rows, cols);
  else if (src_dtype == 1 && dst_dtype == 1)
    transpose_bcs_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), rows, cols);
  else if (src_dtype == 1 && dst_dtype == 0)
    transpose_bcs_kernel<__nv_bfloat16, float><<<grid, block, 0, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                                       static_cast<float*>(out), rows, cols);
  else if (src_dtype == 0 && dst_dtype == 0)
    transpose_bcs_kernel<float, float><<<grid, block, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(out),
                                                               rows, cols);
  else {
    set_error("transpose: unsupported dtype combination %d -> %d", src_dtype, dst_dtype);
    return MVD_ERR_INVALID;
  }
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_cfg_ddpm_step_f32(const float* model_out, float* latents, const float* noise, int64_t n, int cfg,
                          float guidance, float sqrt_alpha_bar, float sqrt_one_minus_alpha_bar, float coef_x0,
                          float coef_xt, float sigma, void* stream) {
  using namespace mvd;
  MVD_CHECK(n > 0 && (cfg == 1 || cfg == 2), "cfg_ddpm_step: bad arguments");
  cfg_ddpm_step_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      model_out, latents, noise, n, cfg, guidance, sqrt_alpha_bar, sqrt_one_minus_alpha_bar, coef_x0, coef_xt, sigma);
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_cfg_ddpm_step_table_f32(const float* model_out, float* latents, const float* noise_table, int64_t n, int cfg,
                                float guidance, const float* coef_table, const int* step_idx, void* stream) {
  using namespace mvd;
  MVD_CHECK(n > 0 && (cfg == 1 || cfg == 2) && coef_table != nullptr && step_idx != nullptr,
            "cfg_ddpm_step_table: bad arguments");
  MVD_CUDA(launch_pdl(cfg_ddpm_step_table_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0,
                      static_cast<cudaStream_t>(stream), model_out, latents, noise_table, n, cfg, guidance, coef_table,
                      step_idx));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_advance_step(int* step_idx, const float* coef_table, float* timestep_out, int n_steps, void* stream) {
  using namespace mvd;
  MVD_CHECK(step_idx != nullptr && coef_table != nullptr && timestep_out != nullptr && n_steps > 0,
            "advance_step: bad arguments");
  MVD_CUDA(launch_pdl(advance_step_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), step_idx, coef_table,
                      timestep_out, n_steps));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

int mvd_advance_step_rows(int* step_idx, const float* coef_table, float* timestep_out, int n_steps,
                          const float* row_table, float* row_out, int row_len, void* stream) {
  using namespace mvd;
  MVD_CHECK(step_idx != nullptr && coef_table != nullptr && timestep_out != nullptr && n_steps > 0 &&
                row_table != nullptr && row_out != nullptr && row_len > 0 && row_len % 4 == 0,
            "advance_step_rows: bad arguments");
  MVD_CHECK(((reinterpret_cast<uintptr_t>(row_table) | reinterpret_cast<uintptr_t>(row_out)) & 15) == 0,
            "advance_step_rows: rows must be 16-byte aligned");
  MVD_CUDA(launch_pdl(advance_step_rows_kernel, dim3(1), dim3(1024), 0, static_cast<cudaStream_t>(stream), step_idx,
                      coef_table, timestep_out, n_steps, reinterpret_cast<const float4*>(row_table),
                      reinterpret_cast<float4*>(row_out), row_len / 4));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

}  // extern "C"
