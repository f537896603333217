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
  if (two_tiles) {
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
