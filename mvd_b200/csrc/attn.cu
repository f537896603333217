// attn.cu — fused flash-attention forward for head_dim 64 on tcgen05 / TMEM / TMA (sm_100a).
//
// Replaces F.scaled_dot_product_attention at src/models/attention.py:148 (the reference-image /
// cross-view branch: S_q = HW, S_kv = any — all N views' tokens concatenated in north-star mode) and the
// SDPA inside diffusers AttnProcessor2_0 that the reference calls as `original_processor`
// (src/models/attention.py:62-70): self-attention (S_kv = HW) and text cross-attention (S_kv = 77).
//
// Operands are read in place from the projection GEMM outputs: Q/K/V are [B, S, ld] bf16 matrices and a
// head is the 64-column slice at h*64, so no head transpose is ever materialised; O is written to the
// same [B, S_q, ld] layout (column h*64), ready to be the A operand of the output projection.
//
// CTA = one 128-row Q tile of one (batch, head). 192 threads:
//   warp 0      TMA producer (Q once, K/V ring of KS stages)
//   warp 1      TMEM owner + single-thread MMA issuer:  S = Q K^T (SS),  O += P V (A = P from TMEM, B = V MN-major)
//   warps 2..5  softmax, one row per thread: S (fp32, TMEM) -> registers -> online softmax with lazy
//               rescaling -> P (bf16) written back over S in TMEM; final O / l -> bf16 -> global.
// S is double buffered in TMEM so QK^T of block j+1 overlaps the softmax of block j.
#include <stdlib.h>
#include "tc.cuh"
#include "host_common.h"
#include "../../include/mvd_b200.h"

namespace mvd {

constexpr int ATT_BM = 128;   // Q rows per CTA
constexpr int ATT_BN = 128;   // KV rows per block
constexpr int ATT_D = 64;     // head dim
constexpr int ATT_KS = 3;     // K/V smem stages
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB
constexpr int ATT_SMEM = (1 + 2 * ATT_KS) * ATT_TILE_BYTES + 1024 + 1024;
constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_O = 256, TM_COLS = 512;
constexpr float LAZY_RESCALE_THRESHOLD = 8.0f;  // log2 units: P stays below 2^8, exact in fp32/bf16 range

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnArgs {
  int Sq, Skv;
  float scale_log2;  // softmax scale * log2(e)
  __nv_bfloat16* out;
  int64_t ldo;            // elements between consecutive rows of O
  int64_t o_batch_stride; // elements between batches of O
};

__global__ void __launch_bounds__(192, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                const __grid_constant__ CUtensorMap mapV, const AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_KS * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_KS * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* k_full = bars + 1;             // [KS]
  uint64_t* k_empty = k_full + ATT_KS;     // [KS]
  uint64_t* v_full = k_empty + ATT_KS;     // [KS]
  uint64_t* v_empty = v_full + ATT_KS;     // [KS]
  uint64_t* s_full = v_empty + ATT_KS;     // [2]
  uint64_t* p_full = s_full + 2;           // [2]
  uint64_t* o_done = p_full + 2;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_blocks = (p.Skv + ATT_BN - 1) / ATT_BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT_KS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
    }
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_3d(sQ, &mapQ, q_full, head * ATT_D, q_tile * ATT_BM, batch);
      for (int j = 0; j < n_blocks; ++j) {
        const int s = j % ATT_KS;
        const uint32_t ph = (j / ATT_KS) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], ATT_TILE_BYTES);
        tma_load_3d(sK + s * ATT_TILE_BYTES, &mapK, &k_full[s], head * ATT_D, j * ATT_BN, batch);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
        tma_load_3d(sV + s * ATT_TILE_BYTES, &mapV, &v_full[s], head * ATT_D, j * ATT_BN, batch);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_D, false, /*b_mn_major=*/true);
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      auto issue_qk = [&](int j) {
        const int s = j % ATT_KS;
        mbar_wait(&k_full[s], (j / ATT_KS) & 1);
        tc_fence_after();
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + s * ATT_TILE_BYTES));
        const uint32_t d = tmem_base + ((j & 1) ? TM_S1 : TM_S0);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < n_blocks; ++j) {
        if (j + 1 < n_blocks) issue_qk(j + 1);
        const int s = j % ATT_KS;
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        mbar_wait(&v_full[s], (j / ATT_KS) & 1);
        tc_fence_after();
        const uint64_t vdesc = umma_desc_sw128(smem_u32(sV + s * ATT_TILE_BYTES));
        const uint32_t a_tmem = tmem_base + ((j & 1) ? TM_S1 : TM_S0);
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k) {
          // A: 16 bf16 of P per row = 8 TMEM columns; B: 16 kv rows of V = 2048 B (desc units of 16 B)
          umma_ts(tmem_base + TM_O, a_tmem + 8 * k, vdesc + 128 * k, idesc_pv, (j | k) != 0);
        }
        umma_commit(&v_empty[s]);
        umma_commit(o_done);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue =====================
    const int q = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int row_in_tile = q * 32 + lane;
    float m_ref = -INFINITY;  // reference maximum (scaled, log2 domain) the exponentials are relative to
    float l = 0.f;            // running sum of exp2(x - m_ref)

    for (int j = 0; j < n_blocks; ++j) {
      const uint32_t t_s = tmem_base + ((j & 1) ? TM_S1 : TM_S0) + lane_off;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      float x[ATT_BN];
      {
        uint32_t* xr = reinterpret_cast<uint32_t*>(x);
        tmem_ld_32x32b_x32(t_s + 0, xr + 0);
        tmem_ld_32x32b_x32(t_s + 32, xr + 32);
        tmem_ld_32x32b_x32(t_s + 64, xr + 64);
        tmem_ld_32x32b_x32(t_s + 96, xr + 96);
        tmem_ld_wait();
      }
      const int valid = p.Skv - j * ATT_BN;  // columns >= valid are padding (last block only)
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      if (valid >= ATT_BN) {
#pragma unroll
        for (int c = 0; c < ATT_BN; c += 4) {
          mx0 = fmaxf(mx0, x[c]);
          mx1 = fmaxf(mx1, x[c + 1]);
          mx2 = fmaxf(mx2, x[c + 2]);
          mx3 = fmaxf(mx3, x[c + 3]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < ATT_BN; ++c) {
          x[c] = (c < valid) ? x[c] : -INFINITY;
          mx0 = fmaxf(mx0, x[c]);
        }
      }
      // scores are kept raw; the softmax scale (> 0) is folded into one FFMA per element below
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
      // lazy rescale: move the reference only when the maximum grew by more than the threshold
      const bool need = mx > m_ref + LAZY_RESCALE_THRESHOLD;
      float alpha = 1.f;
      if (need) {
        alpha = ex2_approx(m_ref - mx);  // 0 on the first block (m_ref = -inf)
        m_ref = mx;
      }
      const float2 neg_m2 = make_float2(-m_ref, -m_ref);
      const float2 scale2 = make_float2(p.scale_log2, p.scale_log2);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
      uint32_t pk[ATT_BN / 2];
#pragma unroll
      for (int c = 0; c < ATT_BN; c += 4) {  // packed FFMA2 / FADD2: 2 elements per FMA-pipe instruction
        const float2 a0 = ffma2(make_float2(x[c], x[c + 1]), scale2, neg_m2);
        const float2 a1 = ffma2(make_float2(x[c + 2], x[c + 3]), scale2, neg_m2);
        const float2 e0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y));
        const float2 e1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
        acc0 = fadd2(acc0, e0);
        acc1 = fadd2(acc1, e1);
        pk[c >> 1] = pack_bf16x2(e0.x, e0.y);
        pk[(c >> 1) + 1] = pack_bf16x2(e1.x, e1.y);
      }
      const float sum = (acc0.x + acc0.y) + (acc1.x + acc1.y);
      l = l * alpha + sum;

      // O must not be touched (and P(j) aliases nothing PV(j-1) still reads) before PV(j-1) has finished
      if (j > 0) {
        mbar_wait(o_done, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
          const uint32_t t_o = tmem_base + TM_O + lane_off;
          uint32_t o[ATT_D];
          tmem_ld_32x32b_x32(t_o, o);
          tmem_ld_32x32b_x32(t_o + 32, o + 32);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < ATT_D; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
          tmem_st_32x32b_x32(t_o, o);
          tmem_st_32x32b_x32(t_o + 32, o + 32);
        }
      }
      tmem_st_32x32b_x32(t_s, pk);
      tmem_st_32x32b_x32(t_s + 32, pk + 32);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
    }

    // ---- epilogue: O / l -> bf16 -> global
    mbar_wait(o_done, (n_blocks - 1) & 1);
    tc_fence_after();
    {
      const uint32_t t_o = tmem_base + TM_O + lane_off;
      uint32_t o[ATT_D];
      tmem_ld_32x32b_x32(t_o, o);
      tmem_ld_32x32b_x32(t_o + 32, o + 32);
      tmem_ld_wait();
      const float inv_l = 1.f / l;
      const int row = q_tile * ATT_BM + row_in_tile;
      if (row < p.Sq) {
        __nv_bfloat16* dst = p.out + static_cast<int64_t>(batch) * p.o_batch_stride +
                             static_cast<int64_t>(row) * p.ldo + head * ATT_D;
#pragma unroll
        for (int c = 0; c < ATT_D; c += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[c + 0]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}


// ================================================================================================
// Two-tile variant (long sequences): one CTA owns TWO 128-row Q tiles of the same (batch, head) and two
// softmax warpgroups. While warpgroup A runs its exponentials on S_A(j), the tensor pipe computes S_B(j) /
// O_B += P_B V and vice versa, so the MUFU pipe (the bound for head_dim 64) never waits for the tensor pipe and
// every K/V tile fetched by TMA is used by 256 query rows.
//   TMEM: S_A [0,128) S_B [128,256) P_A [256,320) P_B [320,384) O_A [384,448) O_B [448,512)
//   P has its OWN columns: as soon as a warpgroup has pulled S_t(j) into registers it releases S_t (s_free) and
//   the tensor pipe refills it with S_t(j+1) while the warpgroup is still exponentiating block j — the registers
//   act as the second S buffer, which TMEM (512 columns) has no room for at two tiles x 128 KV columns.
//   warp 0 TMA, warp 1 MMA, warps 2..5 softmax A, warps 6..9 softmax B  (320 threads)
// MMA issue order per KV block j:  QK_A(j+1)  QK_B(j+1)  PV_A(j)  PV_B(j)   (prologue: QK_A(0) QK_B(0))
// ================================================================================================
constexpr int ATT2_KS = 3;
constexpr int ATT2_SMEM = (2 + 2 * ATT2_KS) * ATT_TILE_BYTES + 1024 + 1024;
__device__ __forceinline__ constexpr uint32_t tm2_s(int t) { return t ? 128u : 0u; }
__device__ __forceinline__ constexpr uint32_t tm2_p(int t) { return t ? 320u : 256u; }
__device__ __forceinline__ constexpr uint32_t tm2_o(int t) { return t ? 448u : 384u; }

__global__ void __launch_bounds__(320, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;  // two tiles
  uint8_t* sK = smem + 2 * ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT2_KS * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT2_KS * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                   // [1]
  uint64_t* k_full = bars + 1;               // [KS]
  uint64_t* k_empty = k_full + ATT2_KS;      // [KS]
  uint64_t* v_full = k_empty + ATT2_KS;      // [KS]
  uint64_t* v_empty = v_full + ATT2_KS;      // [KS]
  uint64_t* s_full = v_empty + ATT2_KS;      // [2] per tile
  uint64_t* p_full = s_full + 2;             // [2] per tile
  uint64_t* o_final = p_full + 2;            // [1]
  uint64_t* s_free = o_final + 1;            // [2] per tile: S_t has been read into registers
  uint64_t* pv_done = s_free + 2;            // [2] per tile: PV_t(j) complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform role index
  const int lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x;  // rows [256 * q_pair, 256 * q_pair + 256)
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_blocks = (p.Skv + ATT_BN - 1) / ATT_BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT2_KS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);
      mbar_init(&s_free[t], 4);
      mbar_init(&pv_done[t], 1);
    }
    mbar_init(o_final, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
      tma_load_3d(sQ, &mapQ, q_full, head * ATT_D, q_pair * 256, batch);
      tma_load_3d(sQ + ATT_TILE_BYTES, &mapQ, q_full, head * ATT_D, q_pair * 256 + 128, batch);
      for (int j = 0; j < n_blocks; ++j) {
        const int s = j % ATT2_KS;
        const uint32_t ph = (j / ATT2_KS) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], ATT_TILE_BYTES);
        tma_load_3d(sK + s * ATT_TILE_BYTES, &mapK, &k_full[s], head * ATT_D, j * ATT_BN, batch);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
        tma_load_3d(sV + s * ATT_TILE_BYTES, &mapV, &v_full[s], head * ATT_D, j * ATT_BN, batch);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Warp-uniform control flow (all 32 lanes wait and compute descriptors, so they live in uniform registers and
    // the tcgen05 operands need no per-issue R2UR traffic); one elected lane issues the MMAs and commits.
    constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, false, false);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_D, false, /*b_mn_major=*/true);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t q_addr = __shfl_sync(0xffffffffu, smem_u32(sQ), 0);
    const uint32_t k_addr = __shfl_sync(0xffffffffu, smem_u32(sK), 0);
    const uint32_t v_addr = __shfl_sync(0xffffffffu, smem_u32(sV), 0);
    auto issue_qk = [&](int t, int j) {  // S_t = Q_t K_j^T ; K_j has landed (caller waited)
      const uint64_t qdesc = umma_desc_sw128(q_addr + t * ATT_TILE_BYTES);
      const uint64_t kdesc = umma_desc_sw128(k_addr + (j % ATT2_KS) * ATT_TILE_BYTES);
      const uint32_t d = tb + tm2_s(t);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int j) {  // O_t += P_t V_j
      const uint64_t vdesc = umma_desc_sw128(v_addr + (j % ATT2_KS) * ATT_TILE_BYTES);
      const uint32_t a_tmem = tb + tm2_p(t);
      const uint32_t d = tb + tm2_o(t);
      if (elect_one()) {
        umma_ts(d, a_tmem, vdesc, idesc_pv, j != 0);
#pragma unroll
        for (int k = 1; k < ATT_BN / 16; ++k) umma_ts(d, a_tmem + 8 * k, vdesc + 128 * k, idesc_pv, 1);
        umma_commit(&pv_done[t]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    issue_qk(0, 0);
    issue_qk(1, 0);
    if (elect_one()) umma_commit(&k_empty[0]);
    __syncwarp();
    for (int j = 0; j < n_blocks; ++j) {
      const int s = j % ATT2_KS;
      const uint32_t par = j & 1;
      if (j + 1 < n_blocks) {
        const int sn = (j + 1) % ATT2_KS;
        mbar_wait(&k_full[sn], ((j + 1) / ATT2_KS) & 1);
        mbar_wait(&s_free[0], par);
        tc_fence_after();
        issue_qk(0, j + 1);
        mbar_wait(&s_free[1], par);
        tc_fence_after();
        issue_qk(1, j + 1);
        if (elect_one()) umma_commit(&k_empty[sn]);
        __syncwarp();
      }
      mbar_wait(&v_full[s], (j / ATT2_KS) & 1);
      mbar_wait(&p_full[0], par);
      tc_fence_after();
      issue_pv(0, j);
      mbar_wait(&p_full[1], par);
      tc_fence_after();
      issue_pv(1, j);
      if (elect_one()) umma_commit(&v_empty[s]);
      __syncwarp();
    }
    if (elect_one()) umma_commit(o_final);
    __syncwarp();
  } else {
    // ===================== softmax warpgroups (warps 2..5: tile A, 6..9: tile B) =====================
    const int t = (warp - 2) >> 2;  // tile / warpgroup index
    const int q = warp & 3;         // TMEM lane quadrant
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int row_in_tile = q * 32 + lane;
    const uint32_t t_s = tmem_base + tm2_s(t) + lane_off;
    const uint32_t t_p = tmem_base + tm2_p(t) + lane_off;
    const uint32_t t_o = tmem_base + tm2_o(t) + lane_off;
    float m_ref = -INFINITY;
    float l = 0.f;

    for (int j = 0; j < n_blocks; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      float x[ATT_BN];
      {
        uint32_t* xr = reinterpret_cast<uint32_t*>(x);
        tmem_ld_32x32b_x32(t_s + 0, xr + 0);
        tmem_ld_32x32b_x32(t_s + 32, xr + 32);
        tmem_ld_32x32b_x32(t_s + 64, xr + 64);
        tmem_ld_32x32b_x32(t_s + 96, xr + 96);
        tmem_ld_wait();
      }
      // S_t now lives in registers: hand the TMEM buffer back so that QK_t(j+1) overlaps this block's softmax
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);
      const int valid = p.Skv - j * ATT_BN;
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      if (valid >= ATT_BN) {
#pragma unroll
        for (int c = 0; c < ATT_BN; c += 4) {
          mx0 = fmaxf(mx0, x[c]);
          mx1 = fmaxf(mx1, x[c + 1]);
          mx2 = fmaxf(mx2, x[c + 2]);
          mx3 = fmaxf(mx3, x[c + 3]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < ATT_BN; ++c) {
          x[c] = (c < valid) ? x[c] : -INFINITY;
          mx0 = fmaxf(mx0, x[c]);
        }
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
      const bool need = mx > m_ref + LAZY_RESCALE_THRESHOLD;
      float alpha = 1.f;
      if (need) {
        alpha = ex2_approx(m_ref - mx);
        m_ref = mx;
      }
      const float2 neg_m2 = make_float2(-m_ref, -m_ref);
      const float2 scale2 = make_float2(p.scale_log2, p.scale_log2);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
      uint32_t pk[ATT_BN / 2];
#pragma unroll
      for (int c = 0; c < ATT_BN; c += 4) {  // packed FFMA2 / FADD2: 2 elements per FMA-pipe instruction
        const float2 a0 = ffma2(make_float2(x[c], x[c + 1]), scale2, neg_m2);
        const float2 a1 = ffma2(make_float2(x[c + 2], x[c + 3]), scale2, neg_m2);
        const float2 e0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y));
        const float2 e1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
        acc0 = fadd2(acc0, e0);
        acc1 = fadd2(acc1, e1);
        pk[c >> 1] = pack_bf16x2(e0.x, e0.y);
        pk[(c >> 1) + 1] = pack_bf16x2(e1.x, e1.y);
      }
      l = l * alpha + ((acc0.x + acc0.y) + (acc1.x + acc1.y));
      // P_t and O_t may only be touched once PV_t(j-1) has completed (it reads P_t and accumulates into O_t);
      // that MMA was issued a whole softmax phase ago, so this wait is normally already satisfied.
      if (j > 0) {
        mbar_wait(&pv_done[t], (j - 1) & 1);
        tc_fence_after();
      }
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        uint32_t o[ATT_D];
        tmem_ld_32x32b_x32(t_o, o);
        tmem_ld_32x32b_x32(t_o + 32, o + 32);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < ATT_D; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
        tmem_st_32x32b_x32(t_o, o);
        tmem_st_32x32b_x32(t_o + 32, o + 32);
      }
      tmem_st_32x32b_x32(t_p, pk);
      tmem_st_32x32b_x32(t_p + 32, pk + 32);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }

    mbar_wait(o_final, 0);
    tc_fence_after();
    {
      uint32_t o[ATT_D];
      tmem_ld_32x32b_x32(t_o, o);
      tmem_ld_32x32b_x32(t_o + 32, o + 32);
      tmem_ld_wait();
      const float inv_l = 1.f / l;
      const int row = q_pair * 256 + t * 128 + row_in_tile;
      if (row < p.Sq) {
        __nv_bfloat16* dst = p.out + static_cast<int64_t>(batch) * p.o_batch_stride +
                             static_cast<int64_t>(row) * p.ldo + head * ATT_D;
#pragma unroll
        for (int c = 0; c < ATT_D; c += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[c + 0]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}


// ================================================================================================
// Third generation of the two-tile kernel. Same tiling, TMEM layout and K/V ring as attn_fwd2_kernel; what changes
// is the schedule of the softmax warps, which is what bounds head_dim 64 (256 tensor FLOP per exponential):
//  * MUFU ping-pong per SM sub-partition. Warp (t, q) of warpgroup t and warp (1-t, q) of the other warpgroup live
//    on the same sub-partition and share its MUFU unit (4 ex2/clk). Their exponential sections are serialised by a
//    pair of mbarriers (seq[t][q]) so that one warp runs its ex2 stream alone at full MUFU rate while its partner
//    does everything that does not need the MUFU (S wait, TMEM load, row max, scale, P store, barriers). In
//    attn_fwd2_kernel both warpgroups ran in lockstep and the MUFU idled about a third of every KV block.
//  * A compile-time share POLY8/8 of the exponentials is evaluated on the FMA pipe instead (round-down range
//    reduction with the 1.5*2^23 magic constant + degree-3 polynomial, packed fp32x2; relative error 8.8e-5, far
//    below the bf16 rounding of P), taking load off the MUFU.
//  * P is stored to TMEM in 32-column chunks as it is produced (registers: x[128] + 16), the row maximum uses the
//    3-input max, and the wait for PV(j-1) (needed before P / O may be overwritten) sits before the exponential
//    section, outside the serialised region.
//  * MMA issue order follows the half-period phase shift between the warpgroups:
//        QK_A(j+1)  PV_B(j-1)  QK_B(j+1)  PV_A(j)
// ================================================================================================
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float2 fadd2_rm(float2 a, float2 b) {
  float2 d;
  asm("add.rm.ftz.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}
// 2^x for x <= ~10 on the FMA / ALU pipes (no MUFU): x = n + f, n = floor(x), f in [0,1);
// 2^f ~ 1 + f (c1 + f (c2 + f c3)) (minimax, rel. err 8.8e-5), exponent patched in with one shift-add.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  const float kMagic = 12582912.f;  // 1.5 * 2^23: the integer part lands in the low mantissa bits
  x.x = fmaxf(x.x, -127.f);
  x.y = fmaxf(x.y, -127.f);
  const float2 t = fadd2_rm(x, make_float2(kMagic, kMagic));                 // round toward -inf
  const float2 n = fadd2(t, make_float2(-kMagic, -kMagic));                  // floor(x), exact
  const float2 f = ffma2(n, make_float2(-1.f, -1.f), x);                     // x - floor(x), exact
  float2 p = ffma2(f, make_float2(0.077119089663028717f, 0.077119089663028717f),
                   make_float2(0.227564394474029541f, 0.227564394474029541f));
  p = ffma2(p, f, make_float2(0.695146143436431885f, 0.695146143436431885f));
  p = ffma2(p, f, make_float2(1.f, 1.f));
  float2 r;
  r.x = __int_as_float((__float_as_int(t.x) << 23) + __float_as_int(p.x));
  r.y = __int_as_float((__float_as_int(t.y) << 23) + __float_as_int(p.y));
  return r;
}

// SEQ: 0 free-running, 1 strict alternation of the exponential sections, 2 tile B delayed once (first block) by one
// section, 3 one-sided (B waits for A every block, A never waits).  SCALE_IN: x*scale - m inside the section.
template <int SEQ, int POLY8, bool SCALE_IN, bool TRACE>
__global__ void __launch_bounds__(384, 1)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const AttnArgs p, long long* trace) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;  // two tiles
  uint8_t* sK = smem + 2 * ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT2_KS * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT2_KS * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                   // [1]
  uint64_t* k_full = bars + 1;               // [KS]
  uint64_t* k_empty = k_full + ATT2_KS;      // [KS]
  uint64_t* v_full = k_empty + ATT2_KS;      // [KS]
  uint64_t* v_empty = v_full + ATT2_KS;      // [KS]
  uint64_t* s_full = v_empty + ATT2_KS;      // [2] per tile
  uint64_t* p_full = s_full + 2;             // [2] per tile
  uint64_t* o_final = p_full + 2;            // [1]
  uint64_t* s_free = o_final + 1;            // [2] per tile: S_t has been read into registers
  uint64_t* pv_done = s_free + 2;            // [2] per tile: PV_t(j) complete
  uint64_t* seq = pv_done + 2;               // [2][4]: warp (t, q) may start its exponential section
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(seq + 8);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_blocks = (p.Skv + ATT_BN - 1) / ATT_BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT2_KS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);
      mbar_init(&s_free[t], 4);
      mbar_init(&pv_done[t], 1);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&seq[i], 1);
    mbar_init(o_final, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
      tma_load_3d(sQ, &mapQ, q_full, head * ATT_D, q_pair * 256, batch);
      tma_load_3d(sQ + ATT_TILE_BYTES, &mapQ, q_full, head * ATT_D, q_pair * 256 + 128, batch);
      for (int j = 0; j < n_blocks; ++j) {
        const int s = j % ATT2_KS;
        const uint32_t ph = (j / ATT2_KS) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], ATT_TILE_BYTES);
        tma_load_3d(sK + s * ATT_TILE_BYTES, &mapK, &k_full[s], head * ATT_D, j * ATT_BN, batch);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
        tma_load_3d(sV + s * ATT_TILE_BYTES, &mapV, &v_full[s], head * ATT_D, j * ATT_BN, batch);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, false, false);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_D, false, /*b_mn_major=*/true);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t q_addr = __shfl_sync(0xffffffffu, smem_u32(sQ), 0);
    const uint32_t k_addr = __shfl_sync(0xffffffffu, smem_u32(sK), 0);
    const uint32_t v_addr = __shfl_sync(0xffffffffu, smem_u32(sV), 0);
    auto issue_qk = [&](int t, int j, bool release_k) {  // S_t = Q_t K_j^T
      const uint64_t qdesc = umma_desc_sw128(q_addr + t * ATT_TILE_BYTES);
      const uint64_t kdesc = umma_desc_sw128(k_addr + (j % ATT2_KS) * ATT_TILE_BYTES);
      const uint32_t d = tb + tm2_s(t);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[t]);
        if (release_k) umma_commit(&k_empty[j % ATT2_KS]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int j, bool release_v) {  // O_t += P_t V_j
      const uint64_t vdesc = umma_desc_sw128(v_addr + (j % ATT2_KS) * ATT_TILE_BYTES);
      const uint32_t a_tmem = tb + tm2_p(t);
      const uint32_t d = tb + tm2_o(t);
      if (elect_one()) {
        umma_ts(d, a_tmem, vdesc, idesc_pv, j != 0);
#pragma unroll
        for (int k = 1; k < ATT_BN / 16; ++k) umma_ts(d, a_tmem + 8 * k, vdesc + 128 * k, idesc_pv, 1);
        umma_commit(&pv_done[t]);
        if (release_v) umma_commit(&v_empty[j % ATT2_KS]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    issue_qk(0, 0, false);
    issue_qk(1, 0, true);
    for (int j = 0; j < n_blocks; ++j) {
      const uint32_t par = j & 1;
      const bool more = j + 1 < n_blocks;
      if (more) {
        mbar_wait(&k_full[(j + 1) % ATT2_KS], ((j + 1) / ATT2_KS) & 1);
        mbar_wait(&s_free[0], par);
        tc_fence_after();
        issue_qk(0, j + 1, false);
      }
      if (j > 0) {
        mbar_wait(&p_full[1], par ^ 1);
        tc_fence_after();
        issue_pv(1, j - 1, true);
      }
      if (more) {
        mbar_wait(&s_free[1], par);
        tc_fence_after();
        issue_qk(1, j + 1, true);
      }
      mbar_wait(&v_full[j % ATT2_KS], (j / ATT2_KS) & 1);
      mbar_wait(&p_full[0], par);
      tc_fence_after();
      issue_pv(0, j, false);
    }
    mbar_wait(&p_full[1], (n_blocks - 1) & 1);
    tc_fence_after();
    issue_pv(1, n_blocks - 1, true);
    if (elect_one()) umma_commit(o_final);
    __syncwarp();
  }
  } else {
    // ===================== softmax warpgroups (warps 4..7: tile A, 8..11: tile B) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int t = (warp - 4) >> 2;
    const int q = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int row_in_tile = q * 32 + lane;
    const uint32_t t_s = tmem_base + tm2_s(t) + lane_off;
    const uint32_t t_p = tmem_base + tm2_p(t) + lane_off;
    const uint32_t t_o = tmem_base + tm2_o(t) + lane_off;
    uint64_t* seq_mine = &seq[t * 4 + q];
    uint64_t* seq_other = &seq[(1 - t) * 4 + q];
    const bool tracer = TRACE && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && q == 0 && lane == 0;
    long long* tr = trace + t * 8 * 64;
    float m_ref = -INFINITY;
    float l = 0.f;

    for (int j = 0; j < n_blocks; ++j) {
      if (tracer && j < 64) tr[j * 8 + 0] = clock64();
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      if (tracer && j < 64) tr[j * 8 + 1] = clock64();
      float x[ATT_BN];
      {
        uint32_t* xr = reinterpret_cast<uint32_t*>(x);
        tmem_ld_32x32b_x32(t_s + 0, xr + 0);
        tmem_ld_32x32b_x32(t_s + 32, xr + 32);
        tmem_ld_32x32b_x32(t_s + 64, xr + 64);
        tmem_ld_32x32b_x32(t_s + 96, xr + 96);
        // P_t and O_t may only be touched once PV_t(j-1) has completed; it was issued a whole period ago, and the
        // barrier round trip hides behind the TMEM loads in flight
        if (j > 0) mbar_wait(&pv_done[t], (j - 1) & 1);
        tmem_ld_wait();
        tc_fence_after();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);
      if (tracer && j < 64) tr[j * 8 + 2] = clock64();
      const int valid = p.Skv - j * ATT_BN;
      float mx0 = -INFINITY, mx1 = -INFINITY;
      if (valid >= ATT_BN) {
        float mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < ATT_BN; c += 8) {
          mx0 = max3f(mx0, x[c], x[c + 1]);
          mx1 = max3f(mx1, x[c + 2], x[c + 3]);
          mx2 = max3f(mx2, x[c + 4], x[c + 5]);
          mx3 = max3f(mx3, x[c + 6], x[c + 7]);
        }
        mx0 = fmaxf(mx0, mx2);
        mx1 = fmaxf(mx1, mx3);
      } else {
#pragma unroll
        for (int c = 0; c < ATT_BN; ++c) {
          x[c] = (c < valid) ? x[c] : -INFINITY;
          mx0 = fmaxf(mx0, x[c]);
        }
      }
      const float mx = fmaxf(mx0, mx1) * p.scale_log2;
      const bool need = mx > m_ref + LAZY_RESCALE_THRESHOLD;
      float alpha = 1.f;
      if (need) {
        alpha = ex2_approx(m_ref - mx);
        m_ref = mx;
      }
      const float2 neg_m2 = make_float2(-m_ref, -m_ref);
      const float2 scale2 = make_float2(p.scale_log2, p.scale_log2);
      float2* x2 = reinterpret_cast<float2*>(x);
      if (!SCALE_IN) {
#pragma unroll
        for (int i = 0; i < ATT_BN / 2; ++i) x2[i] = ffma2(x2[i], scale2, neg_m2);
      }
      if (j > 0) {
        if (__any_sync(0xffffffffu, need)) {
          uint32_t o[ATT_D];
          tmem_ld_32x32b_x32(t_o, o);
          tmem_ld_32x32b_x32(t_o + 32, o + 32);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < ATT_D; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
          tmem_st_32x32b_x32(t_o, o);
          tmem_st_32x32b_x32(t_o + 32, o + 32);
        }
      }
      if (tracer && j < 64) tr[j * 8 + 3] = clock64();
      if (SEQ == 1 && (t == 1 || j > 0)) mbar_wait(seq_mine, (t == 1 ? j : j - 1) & 1);
      if (SEQ == 2 && t == 1 && j == 0) mbar_wait(seq_mine, 0);
      if (SEQ == 3 && t == 1) mbar_wait(seq_mine, j & 1);
      if (tracer && j < 64) tr[j * 8 + 4] = clock64();
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {  // 32 columns -> 16 packed registers -> one TMEM store
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const int i0 = ch * 16 + i;  // pair index 0..63
          float2 a0 = x2[i0], a1 = x2[i0 + 1], e0, e1;
          if (SCALE_IN) {
            a0 = ffma2(a0, scale2, neg_m2);
            a1 = ffma2(a1, scale2, neg_m2);
          }
          if ((i0 & 7) < POLY8) e0 = ex2_poly2(a0);
          else e0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y));
          if (((i0 + 1) & 7) < POLY8) e1 = ex2_poly2(a1);
          else e1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
          acc0 = fadd2(acc0, e0);
          acc1 = fadd2(acc1, e1);
          pk[i] = pack_bf16x2(e0.x, e0.y);
          pk[i + 1] = pack_bf16x2(e1.x, e1.y);
        }
        tmem_st_32x32b_x16(t_p + ch * 16, pk);
      }
      if (SEQ == 1 || (SEQ == 2 && t == 0 && j == 0) || (SEQ == 3 && t == 0)) {
        __syncwarp();
        if (lane == 0) mbar_arrive(seq_other);
      }
      if (tracer && j < 64) tr[j * 8 + 5] = clock64();
      l = l * alpha + ((acc0.x + acc0.y) + (acc1.x + acc1.y));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      if (tracer && j < 64) tr[j * 8 + 6] = clock64();
    }

    mbar_wait(o_final, 0);
    tc_fence_after();
    {
      uint32_t o[ATT_D];
      tmem_ld_32x32b_x32(t_o, o);
      tmem_ld_32x32b_x32(t_o + 32, o + 32);
      tmem_ld_wait();
      const float inv_l = 1.f / l;
      const int row = q_pair * 256 + t * 128 + row_in_tile;
      if (row < p.Sq) {
        __nv_bfloat16* dst = p.out + static_cast<int64_t>(batch) * p.o_batch_stride +
                             static_cast<int64_t>(row) * p.ldo + head * ATT_D;
#pragma unroll
        for (int c = 0; c < ATT_D; c += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[c + 0]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}


// ================================================================================================
// Fourth generation: FOUR softmax warps per SM sub-partition. Measured on B200 (profiles/micro/softmax_pipe.cu): one
// 32x128 softmax block costs a warp ~1050 cycles when the sub-partition's MUFU (8 cycles per MUFU.EX2 warp
// instruction) is the only limit, but 1233 with two resident warps per sub-partition, because every phase that does
// not feed the MUFU (S wait, TMEM load, row max, barriers) of one warp can only be covered by ONE partner. Here each
// 128-row tile is handled by two warpgroups that split the 128 KV columns of a block in halves (64 each): 16 softmax
// warps, four per sub-partition, so three partners cover a warp's non-MUFU phases. The two warps that share rows
// agree on the running reference maximum through shared memory and a 64-thread named barrier per block.
//   threads: warp 0 TMA, warp 1 MMA, warps 2,3 idle (register donors), warps 4..19 softmax:
//            sw = warp - 4: quadrant = sw & 3, tile = (sw >> 2) & 1, column half = sw >> 3
// TMEM layout, K/V ring and MMA issue order are those of attn_fwd3_kernel.
// ================================================================================================
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int POLY8, bool TRACE>
__global__ void __launch_bounds__(640, 1)
attn_fwd4_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const AttnArgs p, long long* trace) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;  // two tiles
  uint8_t* sK = smem + 2 * ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT2_KS * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT2_KS * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                   // [1]
  uint64_t* k_full = bars + 1;               // [KS]
  uint64_t* k_empty = k_full + ATT2_KS;      // [KS]
  uint64_t* v_full = k_empty + ATT2_KS;      // [KS]
  uint64_t* v_empty = v_full + ATT2_KS;      // [KS]
  uint64_t* s_full = v_empty + ATT2_KS;      // [2] per tile
  uint64_t* p_full = s_full + 2;             // [2] per tile (8 warps arrive)
  uint64_t* o_final = p_full + 2;            // [1]
  uint64_t* s_free = o_final + 1;            // [2] per tile (8 warps arrive)
  uint64_t* pv_done = s_free + 2;            // [2] per tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* xch = reinterpret_cast<float*>(tmem_slot + 4);  // [parity 2][tile 2][half 2][128 rows] row-max / row-sum exchange

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_blocks = (p.Skv + ATT_BN - 1) / ATT_BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT2_KS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 8);
      mbar_init(&s_free[t], 8);
      mbar_init(&pv_done[t], 1);
    }
    mbar_init(o_final, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
        tma_load_3d(sQ, &mapQ, q_full, head * ATT_D, q_pair * 256, batch);
        tma_load_3d(sQ + ATT_TILE_BYTES, &mapQ, q_full, head * ATT_D, q_pair * 256 + 128, batch);
        for (int j = 0; j < n_blocks; ++j) {
          const int s = j % ATT2_KS;
          const uint32_t ph = (j / ATT2_KS) & 1;
          mbar_wait(&k_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[s], ATT_TILE_BYTES);
          tma_load_3d(sK + s * ATT_TILE_BYTES, &mapK, &k_full[s], head * ATT_D, j * ATT_BN, batch);
          mbar_wait(&v_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
          tma_load_3d(sV + s * ATT_TILE_BYTES, &mapV, &v_full[s], head * ATT_D, j * ATT_BN, batch);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_D, false, /*b_mn_major=*/true);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t q_addr = __shfl_sync(0xffffffffu, smem_u32(sQ), 0);
      const uint32_t k_addr = __shfl_sync(0xffffffffu, smem_u32(sK), 0);
      const uint32_t v_addr = __shfl_sync(0xffffffffu, smem_u32(sV), 0);
      auto issue_qk = [&](int t, int j, bool release_k) {  // S_t = Q_t K_j^T
        const uint64_t qdesc = umma_desc_sw128(q_addr + t * ATT_TILE_BYTES);
        const uint64_t kdesc = umma_desc_sw128(k_addr + (j % ATT2_KS) * ATT_TILE_BYTES);
        const uint32_t d = tb + tm2_s(t);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
          umma_commit(&s_full[t]);
          if (release_k) umma_commit(&k_empty[j % ATT2_KS]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, int j, bool release_v) {  // O_t += P_t V_j
        const uint64_t vdesc = umma_desc_sw128(v_addr + (j % ATT2_KS) * ATT_TILE_BYTES);
        const uint32_t a_tmem = tb + tm2_p(t);
        const uint32_t d = tb + tm2_o(t);
        if (elect_one()) {
          umma_ts(d, a_tmem, vdesc, idesc_pv, j != 0);
#pragma unroll
          for (int k = 1; k < ATT_BN / 16; ++k) umma_ts(d, a_tmem + 8 * k, vdesc + 128 * k, idesc_pv, 1);
          umma_commit(&pv_done[t]);
          if (release_v) umma_commit(&v_empty[j % ATT2_KS]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0, false);
      issue_qk(1, 0, true);
      for (int j = 0; j < n_blocks; ++j) {
        const uint32_t par = j & 1;
        const bool more = j + 1 < n_blocks;
        if (more) {
          mbar_wait(&k_full[(j + 1) % ATT2_KS], ((j + 1) / ATT2_KS) & 1);
          mbar_wait(&s_free[0], par);
          tc_fence_after();
          issue_qk(0, j + 1, false);
        }
        if (j > 0) {
          mbar_wait(&p_full[1], par ^ 1);
          tc_fence_after();
          issue_pv(1, j - 1, true);
        }
        if (more) {
          mbar_wait(&s_free[1], par);
          tc_fence_after();
          issue_qk(1, j + 1, true);
        }
        mbar_wait(&v_full[j % ATT2_KS], (j / ATT2_KS) & 1);
        mbar_wait(&p_full[0], par);
        tc_fence_after();
        issue_pv(0, j, false);
      }
      mbar_wait(&p_full[1], (n_blocks - 1) & 1);
      tc_fence_after();
      issue_pv(1, n_blocks - 1, true);
      if (elect_one()) umma_commit(o_final);
      __syncwarp();
    }
  } else {
    // ===================== softmax warps: (tile t, column half hf, lane quadrant q) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int sw = warp - 4;
    const int q = sw & 3;
    const int t = (sw >> 2) & 1;
    const int hf = sw >> 3;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int row_in_tile = q * 32 + lane;
    const uint32_t t_s = tmem_base + tm2_s(t) + hf * 64 + lane_off;
    const uint32_t t_p = tmem_base + tm2_p(t) + hf * 32 + lane_off;
    const uint32_t t_o = tmem_base + tm2_o(t) + lane_off;
    const int bar_id = 1 + t * 4 + q;  // named barrier shared with the warp that owns the other column half
    float* xch_mine = xch + (t * 2 + hf) * 128 + row_in_tile;
    float* xch_other = xch + (t * 2 + (1 - hf)) * 128 + row_in_tile;
    const bool tracer = TRACE && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && q == 0 && lane == 0;
    long long* tr = trace + (t * 2 + hf) * 8 * 64;
    float m_ref = -INFINITY;
    float l = 0.f;

    for (int j = 0; j < n_blocks; ++j) {
      if (tracer && j < 64) tr[j * 8 + 0] = clock64();
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      if (tracer && j < 64) tr[j * 8 + 1] = clock64();
      float x[64];
      {
        uint32_t* xr = reinterpret_cast<uint32_t*>(x);
        tmem_ld_32x32b_x32(t_s + 0, xr + 0);
        tmem_ld_32x32b_x32(t_s + 32, xr + 32);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);
      if (tracer && j < 64) tr[j * 8 + 2] = clock64();
      const int valid = p.Skv - j * ATT_BN - hf * 64;  // valid columns of this half (last block only)
      float mx0 = -INFINITY, mx1 = -INFINITY;
      if (valid >= 64) {
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          mx0 = max3f(mx0, x[c], x[c + 1]);
          mx1 = max3f(mx1, x[c + 2], x[c + 3]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          x[c] = (c < valid) ? x[c] : -INFINITY;
          mx0 = fmaxf(mx0, x[c]);
        }
      }
      // agree on the block maximum of the full 128 columns with the warp that owns the other half of these rows
      float* slot_mine = xch_mine + (j & 1) * 512;
      *slot_mine = fmaxf(mx0, mx1);
      named_bar_sync(bar_id, 64);
      const float mx = fmaxf(fmaxf(mx0, mx1), xch_other[(j & 1) * 512]) * p.scale_log2;
      if (tracer && j < 64) tr[j * 8 + 3] = clock64();
      const bool need = mx > m_ref + LAZY_RESCALE_THRESHOLD;
      float alpha = 1.f;
      if (need) {
        alpha = ex2_approx(m_ref - mx);
        m_ref = mx;
      }
      // P_t and O_t may only be touched once PV_t(j-1) has completed
      if (j > 0) {
        mbar_wait(&pv_done[t], (j - 1) & 1);
        tc_fence_after();
        if (hf == 0 && __any_sync(0xffffffffu, need)) {  // rare (lazy rescaling): 8 columns at a time, few registers
#pragma unroll 1
          for (int c0 = 0; c0 < ATT_D; c0 += 8) {
            uint32_t o[8];
            tmem_ld_32x32b_x8(t_o + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            tmem_st_32x32b_x8(t_o + c0, o);
          }
        }
      }
      if (tracer && j < 64) tr[j * 8 + 4] = clock64();
      const float2 neg_m2 = make_float2(-m_ref, -m_ref);
      const float2 scale2 = make_float2(p.scale_log2, p.scale_log2);
      float2* x2 = reinterpret_cast<float2*>(x);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {  // 32 columns -> 16 packed registers -> one TMEM store
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const int i0 = ch * 16 + i;  // pair index 0..31
          float2 a0 = ffma2(x2[i0], scale2, neg_m2), a1 = ffma2(x2[i0 + 1], scale2, neg_m2), e0, e1;
          if ((i0 & 7) < POLY8) e0 = ex2_poly2(a0);
          else e0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y));
          if (((i0 + 1) & 7) < POLY8) e1 = ex2_poly2(a1);
          else e1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
          acc0 = fadd2(acc0, e0);
          acc1 = fadd2(acc1, e1);
          pk[i] = pack_bf16x2(e0.x, e0.y);
          pk[i + 1] = pack_bf16x2(e1.x, e1.y);
        }
        tmem_st_32x32b_x16(t_p + ch * 16, pk);
      }
      if (tracer && j < 64) tr[j * 8 + 5] = clock64();
      l = l * alpha + ((acc0.x + acc0.y) + (acc1.x + acc1.y));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      if (tracer && j < 64) tr[j * 8 + 6] = clock64();
    }

    // ---- epilogue: total row sum = own half + partner's half (same reference maximum), then O / l -> bf16 -> global;
    //      each of the two warps stores 32 of the 64 output columns
    xch_mine[(n_blocks & 1) * 512] = l;
    named_bar_sync(bar_id, 64);
    l += xch_other[(n_blocks & 1) * 512];
    mbar_wait(o_final, 0);
    tc_fence_after();
    {
      uint32_t o[32];
      tmem_ld_32x32b_x32(t_o + hf * 32, o);
      tmem_ld_wait();
      const float inv_l = 1.f / l;
      const int row = q_pair * 256 + t * 128 + row_in_tile;
      if (row < p.Sq) {
        __nv_bfloat16* dst = p.out + static_cast<int64_t>(batch) * p.o_batch_stride +
                             static_cast<int64_t>(row) * p.ldo + head * ATT_D + hf * 32;
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[c + 0]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}


// ================================================================================================
// Fifth generation (two tiles, two softmax warpgroups as attn_fwd3_kernel) built on what the round-2 traces showed
// (profiles/r2_attn_variants*.txt): the kernel is bound by the SERIAL chain of one softmax warp per KV block
// (S wait -> TMEM load -> row max -> exponentials -> P store; 2830 of a 2940-cycle period), in which the MUFU only
// runs during the exponentials. Changes:
//  * look-ahead row maximum: while a warp exponentiates block j it streams S(j+1) — already complete in TMEM —
//    through a 32-register window and folds it into the maximum of the NEXT block. At the top of block j+1 the
//    reference maximum is known before S is loaded, so the exponentials start as soon as the registers land; the
//    row-max pass leaves the chain (exact: same maxima as before, S is simply read twice from TMEM).
//  * P is published in two halves (p_half[t][0/1]); the PV MMA is issued as two groups of 4 k-steps, so PV(j) has
//    finished ~1/2 block earlier and the wait for PV(j-1) before the first P store of block j never stalls.
//  * MMA issue order follows the half-period phase shift of the two warpgroups and doubles as the restoring force
//    towards it:  PV_B(j-1).h0  QK_A(j+1)  PV_A(j).h0  PV_B(j-1).h1  QK_B(j+1)  PV_A(j).h1
// ================================================================================================
template <int POLY8, bool TRACE>
__global__ void __launch_bounds__(384, 1)
attn_fwd5_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const AttnArgs p, long long* trace) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;  // two tiles
  uint8_t* sK = smem + 2 * ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT2_KS * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT2_KS * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                   // [1]
  uint64_t* k_full = bars + 1;               // [KS]
  uint64_t* k_empty = k_full + ATT2_KS;      // [KS]
  uint64_t* v_full = k_empty + ATT2_KS;      // [KS]
  uint64_t* v_empty = v_full + ATT2_KS;      // [KS]
  uint64_t* s_full = v_empty + ATT2_KS;      // [2] per tile
  uint64_t* p_half = s_full + 2;             // [2][2] per tile, per half of the KV block
  uint64_t* o_final = p_half + 4;            // [1]
  uint64_t* s_free = o_final + 1;            // [2] per tile: S_t has been read into registers
  uint64_t* pv_done = s_free + 2;            // [2] per tile: PV_t(j) (both halves) complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_blocks = (p.Skv + ATT_BN - 1) / ATT_BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT2_KS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_half[t * 2], 4);
      mbar_init(&p_half[t * 2 + 1], 4);
      mbar_init(&s_free[t], 4);
      mbar_init(&pv_done[t], 1);
    }
    mbar_init(o_final, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
        tma_load_3d(sQ, &mapQ, q_full, head * ATT_D, q_pair * 256, batch);
        tma_load_3d(sQ + ATT_TILE_BYTES, &mapQ, q_full, head * ATT_D, q_pair * 256 + 128, batch);
        for (int j = 0; j < n_blocks; ++j) {
          const int s = j % ATT2_KS;
          const uint32_t ph = (j / ATT2_KS) & 1;
          mbar_wait(&k_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[s], ATT_TILE_BYTES);
          tma_load_3d(sK + s * ATT_TILE_BYTES, &mapK, &k_full[s], head * ATT_D, j * ATT_BN, batch);
          mbar_wait(&v_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
          tma_load_3d(sV + s * ATT_TILE_BYTES, &mapV, &v_full[s], head * ATT_D, j * ATT_BN, batch);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_D, false, /*b_mn_major=*/true);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t q_addr = __shfl_sync(0xffffffffu, smem_u32(sQ), 0);
      const uint32_t k_addr = __shfl_sync(0xffffffffu, smem_u32(sK), 0);
      const uint32_t v_addr = __shfl_sync(0xffffffffu, smem_u32(sV), 0);
      auto issue_qk = [&](int t, int j, bool release_k) {  // S_t = Q_t K_j^T
        const uint64_t qdesc = umma_desc_sw128(q_addr + t * ATT_TILE_BYTES);
        const uint64_t kdesc = umma_desc_sw128(k_addr + (j % ATT2_KS) * ATT_TILE_BYTES);
        const uint32_t d = tb + tm2_s(t);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
          umma_commit(&s_full[t]);
          if (release_k) umma_commit(&k_empty[j % ATT2_KS]);
        }
        __syncwarp();
      };
      // O_t += P_t[:, 64 h : 64 h + 64] V_j[64 h : 64 h + 64, :]   (4 k-steps of 16)
      auto issue_pv = [&](int t, int j, int h, bool release_v) {
        const uint64_t vdesc = umma_desc_sw128(v_addr + (j % ATT2_KS) * ATT_TILE_BYTES);
        const uint32_t a_tmem = tb + tm2_p(t);
        const uint32_t d = tb + tm2_o(t);
        if (elect_one()) {
#pragma unroll
          for (int k = 4 * h; k < 4 * h + 4; ++k)
            umma_ts(d, a_tmem + 8 * k, vdesc + 128 * k, idesc_pv, (j != 0 || k != 0) ? 1u : 0u);
          if (h == 1) umma_commit(&pv_done[t]);
          if (release_v) umma_commit(&v_empty[j % ATT2_KS]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0, false);
      issue_qk(1, 0, true);
      for (int j = 0; j < n_blocks; ++j) {
        const uint32_t par = j & 1;
        const bool more = j + 1 < n_blocks;
        if (j > 0) {
          mbar_wait(&p_half[2], par ^ 1);
          tc_fence_after();
          issue_pv(1, j - 1, 0, false);
        }
        if (more) {
          mbar_wait(&k_full[(j + 1) % ATT2_KS], ((j + 1) / ATT2_KS) & 1);
          mbar_wait(&s_free[0], par);
          tc_fence_after();
          issue_qk(0, j + 1, false);
        }
        mbar_wait(&v_full[j % ATT2_KS], (j / ATT2_KS) & 1);
        mbar_wait(&p_half[0], par);
        tc_fence_after();
        issue_pv(0, j, 0, false);
        if (j > 0) {
          mbar_wait(&p_half[3], par ^ 1);
          tc_fence_after();
          issue_pv(1, j - 1, 1, true);
        }
        if (more) {
          mbar_wait(&s_free[1], par);
          tc_fence_after();
          issue_qk(1, j + 1, true);
        }
        mbar_wait(&p_half[1], par);
        tc_fence_after();
        issue_pv(0, j, 1, false);
      }
      mbar_wait(&p_half[2], (n_blocks - 1) & 1);
      tc_fence_after();
      issue_pv(1, n_blocks - 1, 0, false);
      mbar_wait(&p_half[3], (n_blocks - 1) & 1);
      tc_fence_after();
      issue_pv(1, n_blocks - 1, 1, true);
      if (elect_one()) umma_commit(o_final);
      __syncwarp();
    }
  } else {
    // ===================== softmax warpgroups (warps 4..7: tile A, 8..11: tile B) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int t = (warp - 4) >> 2;
    const int q = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int row_in_tile = q * 32 + lane;
    const uint32_t t_s = tmem_base + tm2_s(t) + lane_off;
    const uint32_t t_p = tmem_base + tm2_p(t) + lane_off;
    const uint32_t t_o = tmem_base + tm2_o(t) + lane_off;
    const bool tracer = TRACE && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && q == 0 && lane == 0;
    long long* tr = trace + t * 8 * 64;
    float m_ref = -INFINITY;
    float l = 0.f;
    float mxn0 = -INFINITY, mxn1 = -INFINITY;  // running maximum of the NEXT block (raw scores)

    // folds 32 raw score columns [c0, c0 + 32) of block jb, held in the window, into the look-ahead maximum
    auto fold_window = [&](const uint32_t* win, int jb, int c0) {
      const int valid = p.Skv - jb * ATT_BN - c0;  // columns >= valid are padding (last block only)
      if (valid >= 32) {
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          mxn0 = max3f(mxn0, __uint_as_float(win[c]), __uint_as_float(win[c + 1]));
          mxn1 = max3f(mxn1, __uint_as_float(win[c + 2]), __uint_as_float(win[c + 3]));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c < valid) mxn0 = fmaxf(mxn0, __uint_as_float(win[c]));
      }
    };

    // prologue: maximum of block 0 (the only one that is not hidden behind exponentials)
    mbar_wait(&s_full[t], 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < ATT_BN; c0 += 32) {
      uint32_t win[32];
      tmem_ld_32x32b_x32(t_s + c0, win);
      tmem_ld_wait();
      fold_window(win, 0, c0);
    }

    for (int j = 0; j < n_blocks; ++j) {
      if (tracer && j < 64) tr[j * 8 + 0] = clock64();
      // S_t(j) is complete: block 0 was waited for above, block j > 0 by the look-ahead of block j - 1
      float x[ATT_BN];
      {
        uint32_t* xr = reinterpret_cast<uint32_t*>(x);
        tmem_ld_32x32b_x32(t_s + 0, xr + 0);
        tmem_ld_32x32b_x32(t_s + 32, xr + 32);
        tmem_ld_32x32b_x32(t_s + 64, xr + 64);
        tmem_ld_32x32b_x32(t_s + 96, xr + 96);
      }
      const float mx = fmaxf(mxn0, mxn1) * p.scale_log2;
      mxn0 = -INFINITY;
      mxn1 = -INFINITY;
      const bool need = mx > m_ref + LAZY_RESCALE_THRESHOLD;
      float alpha = 1.f;
      if (need) {
        alpha = ex2_approx(m_ref - mx);
        m_ref = mx;
      }
      const bool any_need = __any_sync(0xffffffffu, need);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);
      if (tracer && j < 64) tr[j * 8 + 1] = clock64();
      const int valid = p.Skv - j * ATT_BN;
      if (valid < ATT_BN) {
#pragma unroll
        for (int c = 0; c < ATT_BN; ++c) x[c] = (c < valid) ? x[c] : -INFINITY;
      }
      const float2 neg_m2 = make_float2(-m_ref, -m_ref);
      const float2 scale2 = make_float2(p.scale_log2, p.scale_log2);
      float2* x2 = reinterpret_cast<float2*>(x);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
      const bool look = j + 1 < n_blocks;
      uint32_t win[32];

      // exponentials of 16 score columns [16 * sc, 16 * sc + 16) -> 8 packed registers pk[0..8)
      auto exp16 = [&](int sc, uint32_t* pk) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const int i0 = sc * 8 + i;  // pair index 0..63
          const float2 a0 = ffma2(x2[i0], scale2, neg_m2), a1 = ffma2(x2[i0 + 1], scale2, neg_m2);
          float2 e0, e1;
          if ((i0 & 7) < POLY8) e0 = ex2_poly2(a0);
          else e0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y));
          if (((i0 + 1) & 7) < POLY8) e1 = ex2_poly2(a1);
          else e1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
          acc0 = fadd2(acc0, e0);
          acc1 = fadd2(acc1, e1);
          pk[i] = pack_bf16x2(e0.x, e0.y);
          pk[i + 1] = pack_bf16x2(e1.x, e1.y);
        }
      };

      // ---- first half of the block: columns 0..63
      {
        uint32_t pk[16];
        exp16(0, pk);
        exp16(1, pk + 8);
        // P_t and O_t may only be touched once PV_t(j-1) has completed (its second half was issued at the end of
        // block j-1 and is 4 MMAs long)
        if (j > 0) {
          mbar_wait(&pv_done[t], (j - 1) & 1);
          tc_fence_after();
          if (any_need) {  // rare (lazy rescaling): 8 columns at a time, few registers
#pragma unroll 1
            for (int c0 = 0; c0 < ATT_D; c0 += 8) {
              uint32_t o[8];
              tmem_ld_32x32b_x8(t_o + c0, o);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 8; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
              tmem_st_32x32b_x8(t_o + c0, o);
            }
          }
        }
        tmem_st_32x32b_x16(t_p + 0, pk);
        exp16(2, pk);
        exp16(3, pk + 8);
        tmem_st_32x32b_x16(t_p + 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_half[t * 2]);
      if (tracer && j < 64) tr[j * 8 + 2] = clock64();

      // ---- second half: columns 64..127, interleaved with the look-ahead maximum of block j+1
      if (look) {
        mbar_wait(&s_full[t], (j + 1) & 1);
        tc_fence_after();
        tmem_ld_32x32b_x32(t_s + 0, win);
      }
      if (tracer && j < 64) tr[j * 8 + 3] = clock64();
      {
        uint32_t pk[16];
        exp16(4, pk);
        if (look) {
          tmem_ld_wait();
          fold_window(win, j + 1, 0);
          tmem_ld_32x32b_x32(t_s + 32, win);
        }
        exp16(5, pk + 8);
        tmem_st_32x32b_x16(t_p + 32, pk);
        if (look) {
          tmem_ld_wait();
          fold_window(win, j + 1, 32);
          tmem_ld_32x32b_x32(t_s + 64, win);
        }
        exp16(6, pk);
        if (look) {
          tmem_ld_wait();
          fold_window(win, j + 1, 64);
          tmem_ld_32x32b_x32(t_s + 96, win);
        }
        exp16(7, pk + 8);
        tmem_st_32x32b_x16(t_p + 48, pk);
      }
      if (tracer && j < 64) tr[j * 8 + 4] = clock64();
      l = l * alpha + ((acc0.x + acc0.y) + (acc1.x + acc1.y));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_half[t * 2 + 1]);
      if (look) {
        tmem_ld_wait();
        fold_window(win, j + 1, 96);
      }
      if (tracer && j < 64) tr[j * 8 + 5] = clock64();
    }

    mbar_wait(o_final, 0);
    tc_fence_after();
    {
      uint32_t o[ATT_D];
      tmem_ld_32x32b_x32(t_o, o);
      tmem_ld_32x32b_x32(t_o + 32, o + 32);
      tmem_ld_wait();
      const float inv_l = 1.f / l;
      const int row = q_pair * 256 + t * 128 + row_in_tile;
      if (row < p.Sq) {
        __nv_bfloat16* dst = p.out + static_cast<int64_t>(batch) * p.o_batch_stride +
                             static_cast<int64_t>(row) * p.ldo + head * ATT_D;
#pragma unroll
        for (int c = 0; c < ATT_D; c += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[c + 0]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

}  // namespace mvd

extern "C" int mvd_attention_bf16(const void* q, int64_t ldq, int64_t q_batch_stride, const void* k, int64_t ldk,
                                  int64_t k_batch_stride, const void* v, int64_t ldv, int64_t v_batch_stride,
                                  void* out, int64_t ldo, int64_t o_batch_stride, int batch, int heads, int s_q,
                                  int s_kv, float scale, void* stream) {
  using namespace mvd;
  MVD_CHECK(batch > 0 && heads > 0 && s_q > 0 && s_kv > 0, "attention: empty problem B=%d H=%d Sq=%d Skv=%d", batch,
            heads, s_q, s_kv);
  MVD_CHECK(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && q_batch_stride % 8 == 0 &&
                k_batch_stride % 8 == 0 && v_batch_stride % 8 == 0 && o_batch_stride % 8 == 0,
            "attention: strides must be multiples of 8 elements");
  MVD_CHECK(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
              reinterpret_cast<uintptr_t>(out)) & 15) == 0,
            "attention: pointers must be 16-byte aligned");
  MVD_CHECK(batch <= 65535 && heads <= 65535, "attention: batch/heads exceed grid limits");

  CUtensorMap mQ, mK, mV;
  auto mk = [&](CUtensorMap* m, const void* ptr, int64_t ld, int64_t bstride, int S) -> int {
    const uint64_t dims[3] = {static_cast<uint64_t>(heads) * ATT_D, static_cast<uint64_t>(S),
                              static_cast<uint64_t>(batch)};
    const uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(bstride) * 2};
    const uint32_t box[3] = {ATT_D, 128, 1};
    return make_tmap_bf16(m, ptr, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  if (int e = mk(&mQ, q, ldq, q_batch_stride, s_q)) return e;
  if (int e = mk(&mK, k, ldk, k_batch_stride, s_kv)) return e;
  if (int e = mk(&mV, v, ldv, v_batch_stride, s_kv)) return e;

  AttnArgs a;
  a.Sq = s_q;
  a.Skv = s_kv;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.out = static_cast<__nv_bfloat16*>(out);
  a.ldo = ldo;
  a.o_batch_stride = o_batch_stride;

  static bool configured = false;
  if (!configured) {
    MVD_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    MVD_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM));
    configured = true;
  }
  // long query sequences: two Q tiles per CTA (ping-pong softmax warpgroups); short ones: one tile per CTA so
  // that small sites still spread over the SMs
  const bool two_tiles = (s_q >= 512) && (static_cast<long>((s_q + 255) / 256) * heads * batch >= 2L * sm_count());
  const char* var_env = getenv("MVD_ATTN_VARIANT");
  const int variant = var_env ? atoi(var_env) : 0;
  if (two_tiles && variant >= 5000) {
    dim3 grid((s_q + 255) / 256, heads, batch);
    long long* trace = reinterpret_cast<long long*>(getenv("MVD_ATTN_TRACE_PTR") ? strtoull(getenv("MVD_ATTN_TRACE_PTR"), nullptr, 0) : 0ull);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MVD_A5(POLY, TR)                                                                                             \
  do {                                                                                                               \
    MVD_CUDA(cudaFuncSetAttribute(attn_fwd5_kernel<POLY, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM)); \
    MVD_CUDA(launch_pdl(attn_fwd5_kernel<POLY, TR>, grid, dim3(384), ATT2_SMEM, st, mQ, mK, mV, a, trace));           \
  } while (0)
    const int poly = variant % 10;
    const bool tr = (variant % 1000) >= 100 && trace != nullptr;
    if (tr) { if (poly == 0) MVD_A5(0, true); else MVD_A5(1, true); }
    else if (poly == 0) MVD_A5(0, false);
    else if (poly == 1) MVD_A5(1, false);
    else if (poly == 2) MVD_A5(2, false);
    else MVD_A5(3, false);
#undef MVD_A5
  } else if (two_tiles && variant >= 4000) {
    dim3 grid((s_q + 255) / 256, heads, batch);
    long long* trace = reinterpret_cast<long long*>(getenv("MVD_ATTN_TRACE_PTR") ? strtoull(getenv("MVD_ATTN_TRACE_PTR"), nullptr, 0) : 0ull);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    constexpr int SMEM4 = ATT2_SMEM + 4096;
#define MVD_A4(POLY, TR)                                                                                          \
  do {                                                                                                            \
    MVD_CUDA(cudaFuncSetAttribute(attn_fwd4_kernel<POLY, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM4)); \
    MVD_CUDA(launch_pdl(attn_fwd4_kernel<POLY, TR>, grid, dim3(640), SMEM4, st, mQ, mK, mV, a, trace));            \
  } while (0)
    const int poly = variant % 10;
    const bool tr = (variant % 1000) >= 100 && trace != nullptr;
    if (tr) { if (poly == 0) MVD_A4(0, true); else MVD_A4(1, true); }
    else if (poly == 0) MVD_A4(0, false);
    else if (poly == 1) MVD_A4(1, false);
    else if (poly == 2) MVD_A4(2, false);
    else MVD_A4(3, false);
#undef MVD_A4
  } else if (two_tiles && variant != 0) {
    dim3 grid((s_q + 255) / 256, heads, batch);
    long long* trace = reinterpret_cast<long long*>(getenv("MVD_ATTN_TRACE_PTR") ? strtoull(getenv("MVD_ATTN_TRACE_PTR"), nullptr, 0) : 0ull);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MVD_A3(SEQ, POLY, SC, TR)                                                                                    \
  do {                                                                                                               \
    MVD_CUDA(cudaFuncSetAttribute(attn_fwd3_kernel<SEQ, POLY, SC, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                  ATT2_SMEM));                                                                       \
    MVD_CUDA(launch_pdl(attn_fwd3_kernel<SEQ, POLY, SC, TR>, grid, dim3(384), ATT2_SMEM, st, mQ, mK, mV, a, trace)); \
  } while (0)
#define MVD_A3_POLY(SEQ, SC, TR)                                                    \
  do {                                                                              \
    if (poly == 0) MVD_A3(SEQ, 0, SC, TR);                                          \
    else if (poly == 2) MVD_A3(SEQ, 2, SC, TR);                                     \
    else MVD_A3(SEQ, 3, SC, TR);                                                    \
  } while (0)
#define MVD_A3_SEQ(SC, TR)                                                          \
  do {                                                                              \
    if (seqv == 0) MVD_A3_POLY(0, SC, TR);                                          \
    else if (seqv == 1) MVD_A3_POLY(1, SC, TR);                                     \
    else if (seqv == 2) MVD_A3_POLY(2, SC, TR);                                     \
    else MVD_A3_POLY(3, SC, TR);                                                    \
  } while (0)
    // variant = 1 + poly8 + 10*seq + 100*scale_in + 1000*trace
    const int v = variant - 1;
    const bool tr = v >= 1000 && trace != nullptr;
    const int seqv = (v % 100) / 10, poly = v % 10, sc = (v % 1000) / 100;
    if (tr) { if (sc) MVD_A3_SEQ(true, true); else MVD_A3_SEQ(false, true); }
    else { if (sc) MVD_A3_SEQ(true, false); else MVD_A3_SEQ(false, false); }
#undef MVD_A3_SEQ
#undef MVD_A3_POLY
#undef MVD_A3
  } else if (two_tiles) {
    dim3 grid((s_q + 255) / 256, heads, batch);
    MVD_CUDA(launch_pdl(attn_fwd2_kernel, grid, dim3(320), ATT2_SMEM, static_cast<cudaStream_t>(stream), mQ, mK, mV, a));
  } else {
    dim3 grid((s_q + ATT_BM - 1) / ATT_BM, heads, batch);
    MVD_CUDA(launch_pdl(attn_fwd_kernel, grid, dim3(192), ATT_SMEM, static_cast<cudaStream_t>(stream), mQ, mK, mV, a));
  }
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}
