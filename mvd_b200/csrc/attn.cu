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
  int kv_batch_mul;       // 1, or 0 when K/V are shared by every batch entry (batch coordinate 0)
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
        tma_load_3d(sK + s * ATT_TILE_BYTES, &mapK, &k_full[s], head * ATT_D, j * ATT_BN, batch * p.kv_batch_mul);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
        tma_load_3d(sV + s * ATT_TILE_BYTES, &mapV, &v_full[s], head * ATT_D, j * ATT_BN, batch * p.kv_batch_mul);
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
// Two-tile kernel (long sequences): one CTA owns TWO 128-row Q tiles of one (batch, head) and two softmax warpgroups,
// so every K/V tile fetched by TMA serves 256 query rows and the tensor pipe always has the other tile's MMAs to run
// while one warpgroup is in its exponentials.
//   TMEM: S_A [0,128) S_B [128,256) P_A [256,320) P_B [320,384) O_A [384,448) O_B [448,512)
//   P has its OWN columns: as soon as a warpgroup has pulled S_t(j) into registers it releases S_t (s_free) and the
//   tensor pipe refills it with S_t(j+1) while the warpgroup is still exponentiating block j — the registers act as
//   the second S buffer, which TMEM (512 columns) has no room for at two tiles x 128 KV columns.
//   384 threads: warp 0 TMA, warp 1 MMA, warps 2-3 idle (register donors: setmaxnreg 56 / 224), warps 4..7 softmax A,
//   8..11 softmax B. MMA issue order per KV block j:  QK_A(j+1)  PV_B(j-1)  QK_B(j+1)  PV_A(j)  — the order the two
//   warpgroups reach those points once they run half a block apart, which is also where they settle.
//
// What round 2 measured about the softmax side, which bounds head_dim 64 (profiles/r2_attn_variants*.txt,
// profiles/r2_softmax_pipe.txt): the two warps that share an SM sub-partition's MUFU settle in anti-phase by
// themselves and the period of a KV block is the serial chain of ONE warp (S wait, TMEM load, row max, 128 MUFU.EX2
// at 8 cycles each, P store). Strict MUFU hand-over between the two warps (per-sub-partition mbarriers), a one-sided
// hand-over, four warps per sub-partition (column halves, row maximum agreed through shared memory), a look-ahead
// row maximum with split P publication and a two-half software pipeline of the block (TMEM load and row maximum of the
// next half hidden behind the exponentials) were all built and timed: 264-306 us against 250-257 us for this
// free-running form at the configs[1] top site; ncu: XU pipe 67 %, tensor pipe 33 % active (r2_attn_pair_ncu.txt). Packed fp32x2 instructions issue at half rate (no throughput gain over scalar,
// only fewer issue slots), and a degree-3 FMA-pipe exp2 costs ~7.4 issue cycles per element against 8 MUFU cycles,
// so the polynomial share (POLY8 eighths) buys little; it stays a template parameter.
//
// Wave tail: `units` = (256-row pair, head, batch). Whole waves run one unit per CTA; the units of the last, partial
// wave are split along KV into `split` parts each (grid = n_full + rem * split <= one extra wave of short CTAs). A
// part writes its un-normalised fp32 O, reference maximum and row sum to a workspace; the CTA that finishes a unit
// last (ticket counter, re-armed for the next launch) merges the parts in part order — deterministic — and stores
// the bf16 rows. No CTA ever waits for another one.
// ================================================================================================
constexpr int ATT2_KS = 3;
constexpr int ATT2_SMEM = (2 + 2 * ATT2_KS) * ATT_TILE_BYTES + 1024 + 1024;
constexpr int ATT2_THREADS = 384;
constexpr int ATT_MAX_SPLIT = 8;
__device__ __forceinline__ constexpr uint32_t tm2_s(int t) { return t ? 128u : 0u; }
__device__ __forceinline__ constexpr uint32_t tm2_p(int t) { return t ? 320u : 256u; }
__device__ __forceinline__ constexpr uint32_t tm2_o(int t) { return t ? 448u : 384u; }

struct PairSched {
  int q_pairs, heads;   // unit -> (q_pair, head, batch), q_pair fastest
  int n_full;           // units [0, n_full) are processed whole by CTA `unit`
  int split;            // every further unit is processed by `split` CTAs
  float* ws_o;          // [slots][256][64] fp32 partial O (slot = (unit - n_full) * split + part)
  float* ws_ml;         // [slots][256][2]  reference maximum (log2 domain), row sum
  unsigned int* ws_cnt; // [units - n_full] arrival tickets; zero between launches
};

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


template <int POLY8, bool TRACE>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attn_pair_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const AttnArgs p, const PairSched sc, long long* trace) {
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
  uint32_t* last_flag = tmem_slot + 1;       // split units: this CTA took the last ticket

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform role index
  const int lane = threadIdx.x & 31;
  // ---- work item: a whole unit, or one KV part of a unit of the last wave
  int unit = blockIdx.x, part = 0, nparts = 1;
  if (unit >= sc.n_full) {
    const int r = unit - sc.n_full;
    unit = sc.n_full + r / sc.split;
    part = r % sc.split;
    nparts = sc.split;
  }
  const int q_pair = unit % sc.q_pairs;
  const int head = (unit / sc.q_pairs) % sc.heads;
  const int batch = unit / (sc.q_pairs * sc.heads);
  const int n_blocks_all = (p.Skv + ATT_BN - 1) / ATT_BN;
  const int kb0 = n_blocks_all * part / nparts;              // KV blocks [kb0, kb0 + n_blocks) of this item
  const int n_blocks = n_blocks_all * (part + 1) / nparts - kb0;

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
          tma_load_3d(sK + s * ATT_TILE_BYTES, &mapK, &k_full[s], head * ATT_D, (kb0 + j) * ATT_BN, batch * p.kv_batch_mul);
          mbar_wait(&v_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
          tma_load_3d(sV + s * ATT_TILE_BYTES, &mapV, &v_full[s], head * ATT_D, (kb0 + j) * ATT_BN, batch * p.kv_batch_mul);
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
          // A: 16 bf16 of P per row = 8 TMEM columns; B: 16 kv rows of V = 2048 B (descriptor units of 16 B)
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
    const int t = (warp - 4) >> 2;  // tile / warpgroup index
    const int q = warp & 3;         // TMEM lane quadrant
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int row_in_tile = q * 32 + lane;
    const uint32_t t_s = tmem_base + tm2_s(t) + lane_off;
    const uint32_t t_p = tmem_base + tm2_p(t) + lane_off;
    const uint32_t t_o = tmem_base + tm2_o(t) + lane_off;
    const bool tracer = TRACE && blockIdx.x == 0 && q == 0 && lane == 0;
    long long* tr = trace + t * 8 * 64;
    // CTA-level stamps of the LAST CTA of the grid (a KV part when the tail is split): trace[1024 + i]
    const bool tracer2 = TRACE && blockIdx.x == gridDim.x - 1 && warp == 4 && lane == 0;
    if (tracer2) {
      trace[1024] = clock64();
      trace[1030] = n_blocks;
      trace[1031] = nparts;
    }
    float m_ref = -INFINITY;  // reference maximum (scaled, log2 domain) the exponentials are relative to
    float l = 0.f;            // running sum of exp2(x - m_ref)

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
        // P_t and O_t may only be touched once PV_t(j-1) has completed (it reads P_t and accumulates into O_t). The
        // wait sits HERE, before S_t is handed back, on purpose: PV_B(j-1) is issued behind QK_A(j+1), so in lockstep
        // tile B stalls at this point and releases S_B later than tile A releases S_A; the delay accumulates until
        // the two warpgroups run half a block apart (measured: with the wait after the row maximum they stay in
        // lockstep, both exponential phases share the MUFU and the period grows from 2940 to 3440 cycles).
        if (j > 0) mbar_wait(&pv_done[t], (j - 1) & 1);
        tmem_ld_wait();
        tc_fence_after();
      }
      // S_t now lives in registers: hand the TMEM buffer back so that QK_t(j+1) overlaps this block's softmax
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);
      if (tracer && j < 64) tr[j * 8 + 2] = clock64();
      const int valid = p.Skv - (kb0 + j) * ATT_BN;  // columns >= valid are padding (last block only)
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
      // scores are kept raw; the softmax scale (> 0) is folded into one FFMA per element below
      const float mx = fmaxf(mx0, mx1) * p.scale_log2;
      // lazy rescale: move the reference only when the maximum grew by more than the threshold
      const bool need = mx > m_ref + LAZY_RESCALE_THRESHOLD;
      float alpha = 1.f;
      if (need) {
        alpha = ex2_approx(m_ref - mx);  // 0 on the first block (m_ref = -inf)
        m_ref = mx;
      }
      const float2 neg_m2 = make_float2(-m_ref, -m_ref);
      const float2 scale2 = make_float2(p.scale_log2, p.scale_log2);
      float2* x2 = reinterpret_cast<float2*>(x);
      if (tracer && j < 64) tr[j * 8 + 3] = clock64();
      if (j > 0) {
        if (__any_sync(0xffffffffu, need)) {  // rare: 8 columns at a time, few registers
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
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {  // 32 columns -> 16 packed registers -> one TMEM store
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const int i0 = ch * 16 + i;  // pair index 0..63
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

    // ---- epilogue
    if (tracer2) trace[1025] = clock64();
    mbar_wait(o_final, 0);
    tc_fence_after();
    if (tracer2) trace[1026] = clock64();
    float o[ATT_D];
    {
      uint32_t* orr = reinterpret_cast<uint32_t*>(o);
      tmem_ld_32x32b_x32(t_o, orr);
      tmem_ld_32x32b_x32(t_o + 32, orr + 32);
      tmem_ld_wait();
    }
    const int row_in_unit = t * 128 + row_in_tile;
    bool store = true;
    if (nparts > 1) {
      // publish this part, take a ticket; the CTA that takes the last one merges all parts of the unit
      const int su = unit - sc.n_full;
      // partial O is stored column-chunk major, [slot][16 float4 chunks][256 rows], so that the 32 rows of a warp
      // write (and the merging CTA reads) 512 contiguous bytes per instruction
      const size_t slot = static_cast<size_t>(su) * nparts + part;
      float4* wo = reinterpret_cast<float4*>(sc.ws_o) + slot * (16 * 256) + row_in_unit;
#pragma unroll
      for (int c = 0; c < ATT_D / 4; ++c) wo[c * 256] = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
      reinterpret_cast<float2*>(sc.ws_ml)[slot * 256 + row_in_unit] = make_float2(m_ref, l);
      __threadfence();
      if (tracer2) trace[1027] = clock64();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 128) {
        const unsigned int old = atomicAdd(sc.ws_cnt + su, 1u);
        const bool last = old == static_cast<unsigned int>(nparts - 1);
        if (last) sc.ws_cnt[su] = 0u;  // re-arm for the next launch
        *last_flag = last ? 1u : 0u;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      store = *last_flag != 0u;
      if (store) {
        __threadfence();
        const size_t slot0 = static_cast<size_t>(su) * nparts;
        const float2* mlp = reinterpret_cast<const float2*>(sc.ws_ml) + slot0 * 256 + row_in_unit;
        float2 ml[ATT_MAX_SPLIT];
        float m = -INFINITY;
#pragma unroll
        for (int pp = 0; pp < ATT_MAX_SPLIT; ++pp) {
          if (pp < nparts) {
            ml[pp] = __ldcg(mlp + pp * 256);
            m = fmaxf(m, ml[pp].x);
          }
        }
        l = 0.f;
#pragma unroll
        for (int c = 0; c < ATT_D; ++c) o[c] = 0.f;
#pragma unroll 1
        for (int pp = 0; pp < nparts; ++pp) {  // fixed order: the result does not depend on which CTA came last
          float w = 0.f;
#pragma unroll
          for (int k = 0; k < ATT_MAX_SPLIT; ++k)
            if (k == pp) w = ex2_approx(ml[k].x - m) , l = fmaf(w, ml[k].y, l);
          const float4* src = reinterpret_cast<const float4*>(sc.ws_o) + (slot0 + pp) * (16 * 256) + row_in_unit;
          float4 f[ATT_D / 4];
#pragma unroll
          for (int c = 0; c < ATT_D / 4; ++c) f[c] = __ldcg(src + c * 256);
#pragma unroll
          for (int c = 0; c < ATT_D / 4; ++c) {
            o[4 * c] = fmaf(w, f[c].x, o[4 * c]);
            o[4 * c + 1] = fmaf(w, f[c].y, o[4 * c + 1]);
            o[4 * c + 2] = fmaf(w, f[c].z, o[4 * c + 2]);
            o[4 * c + 3] = fmaf(w, f[c].w, o[4 * c + 3]);
          }
        }
      }
    }
    if (tracer2) trace[1028] = clock64();
    const int row = q_pair * 256 + row_in_unit;
    if (store && row < p.Sq) {
      const float inv_l = 1.f / l;
      __nv_bfloat16* dst = p.out + static_cast<int64_t>(batch) * p.o_batch_stride +
                           static_cast<int64_t>(row) * p.ldo + head * ATT_D;
#pragma unroll
      for (int c = 0; c < ATT_D; c += 8) {
        uint4 v;
        v.x = pack_bf16x2(o[c + 0] * inv_l, o[c + 1] * inv_l);
        v.y = pack_bf16x2(o[c + 2] * inv_l, o[c + 3] * inv_l);
        v.z = pack_bf16x2(o[c + 4] * inv_l, o[c + 5] * inv_l);
        v.w = pack_bf16x2(o[c + 6] * inv_l, o[c + 7] * inv_l);
        *reinterpret_cast<uint4*>(dst + c) = v;
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
// Persistent form of attn_pair_kernel for launches of at least one wave of units: one CTA per SM walks the work items
// (whole units first, then the KV parts of the split tail) round-robin. What it buys over one CTA per item: barrier
// initialisation, TMEM allocation and descriptor prefetch happen once per SM; Q of the next item is prefetched into a
// second shared-memory buffer and K/V keep streaming through the ring across the item boundary; the MMA warp issues
// QK(0) of the next item in the slot where QK(j+1) would go, so S is already waiting when the softmax warps come back
// from the epilogue, whose global stores then overlap the next item's first blocks. All barriers run on global
// counters (g: KV blocks of this CTA so far, it: items so far) instead of per-item ones; two more barriers order the
// item boundary: q_empty (all QK of an item retired -> its Q buffer may be refilled) and o_free (the epilogue has
// read O_t from TMEM -> PV_t(0) of the next item may overwrite it).
// ================================================================================================
constexpr int ATT3_SMEM = (4 + 2 * ATT2_KS + 2) * ATT_TILE_BYTES + 1024 + 1024;  // Q x2, K/V ring, O staging
constexpr int ATT_PERSIST_MAX_BLOCKS = 96;  // KV blocks per item up to which the persistent form is used

struct PairItem {
  int unit, part, nparts, q_pair, head, batch, kb0, n_blocks;
};
__device__ __forceinline__ PairItem pair_item(const PairSched& sc, int skv, int idx) {
  PairItem w;
  w.unit = idx;
  w.part = 0;
  w.nparts = 1;
  if (idx >= sc.n_full) {
    const int r = idx - sc.n_full;
    w.unit = sc.n_full + r / sc.split;
    w.part = r % sc.split;
    w.nparts = sc.split;
  }
  w.q_pair = w.unit % sc.q_pairs;
  w.head = (w.unit / sc.q_pairs) % sc.heads;
  w.batch = w.unit / (sc.q_pairs * sc.heads);
  const int n_all = (skv + ATT_BN - 1) / ATT_BN;
  w.kb0 = n_all * w.part / w.nparts;
  w.n_blocks = n_all * (w.part + 1) / w.nparts - w.kb0;
  return w;
}

template <int POLY8, bool TRACE = false>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attn_pair_persist_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                         const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapO,
                         const AttnArgs p, const PairSched sc, const int n_items, long long* trace = nullptr) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;  // two buffers of two tiles
  uint8_t* sK = smem + 4 * ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT2_KS * ATT_TILE_BYTES;
  uint8_t* sO = sV + ATT2_KS * ATT_TILE_BYTES;  // [2 tiles][4 warps][32 rows x 128 B], 128B-swizzled, for the TMA store
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                   // [2] per Q buffer
  uint64_t* q_empty = bars + 2;              // [2]
  uint64_t* k_full = bars + 4;               // [KS]
  uint64_t* k_empty = k_full + ATT2_KS;      // [KS]
  uint64_t* v_full = k_empty + ATT2_KS;      // [KS]
  uint64_t* v_empty = v_full + ATT2_KS;      // [KS]
  uint64_t* s_full = v_empty + ATT2_KS;      // [2] per tile
  uint64_t* p_full = s_full + 2;             // [2]
  uint64_t* s_free = p_full + 2;             // [2]
  uint64_t* pv_done = s_free + 2;            // [2]
  uint64_t* o_free = pv_done + 2;            // [2]
  uint64_t* o_final = o_free + 2;            // [2] per tile: the two warpgroups must NOT meet at the item boundary
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_final + 2);
  uint32_t* last_flag = tmem_slot + 1;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int n_ctas = gridDim.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    tma_prefetch_desc(&mapO);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], 1);
    }
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
      mbar_init(&o_free[t], 4);
      mbar_init(&o_final[t], 1);
    }
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
        uint32_t g = 0;
        int it = 0;
        for (int idx = blockIdx.x; idx < n_items; idx += n_ctas, ++it) {
          const PairItem w = pair_item(sc, p.Skv, idx);
          const int qb = it & 1;
          mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
          uint8_t* q_dst = sQ + qb * 2 * ATT_TILE_BYTES;
          mbar_arrive_expect_tx(&q_full[qb], 2 * ATT_TILE_BYTES);
          tma_load_3d(q_dst, &mapQ, &q_full[qb], w.head * ATT_D, w.q_pair * 256, w.batch);
          tma_load_3d(q_dst + ATT_TILE_BYTES, &mapQ, &q_full[qb], w.head * ATT_D, w.q_pair * 256 + 128, w.batch);
          for (int j = 0; j < w.n_blocks; ++j, ++g) {
            const int s = g % ATT2_KS;
            const uint32_t ph = (g / ATT2_KS) & 1;
            mbar_wait(&k_empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&k_full[s], ATT_TILE_BYTES);
            tma_load_3d(sK + s * ATT_TILE_BYTES, &mapK, &k_full[s], w.head * ATT_D, (w.kb0 + j) * ATT_BN,
                        w.batch * p.kv_batch_mul);
            mbar_wait(&v_empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
            tma_load_3d(sV + s * ATT_TILE_BYTES, &mapV, &v_full[s], w.head * ATT_D, (w.kb0 + j) * ATT_BN,
                        w.batch * p.kv_batch_mul);
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_D, false, /*b_mn_major=*/true);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t q_addr = __shfl_sync(0xffffffffu, smem_u32(sQ), 0);
      const uint32_t k_addr = __shfl_sync(0xffffffffu, smem_u32(sK), 0);
      const uint32_t v_addr = __shfl_sync(0xffffffffu, smem_u32(sV), 0);
      auto issue_qk = [&](int t, int qb, int slot, bool release_k) {  // S_t = Q_t K^T
        const uint64_t qdesc = umma_desc_sw128(q_addr + (qb * 2 + t) * ATT_TILE_BYTES);
        const uint64_t kdesc = umma_desc_sw128(k_addr + slot * ATT_TILE_BYTES);
        const uint32_t d = tb + tm2_s(t);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
          umma_commit(&s_full[t]);
          if (release_k) umma_commit(&k_empty[slot]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, int slot, bool accumulate, bool release_v) {  // O_t (+)= P_t V
        const uint64_t vdesc = umma_desc_sw128(v_addr + slot * ATT_TILE_BYTES);
        const uint32_t a_tmem = tb + tm2_p(t);
        const uint32_t d = tb + tm2_o(t);
        if (elect_one()) {
          umma_ts(d, a_tmem, vdesc, idesc_pv, accumulate);
#pragma unroll
          for (int k = 1; k < ATT_BN / 16; ++k) umma_ts(d, a_tmem + 8 * k, vdesc + 128 * k, idesc_pv, 1);
          umma_commit(&pv_done[t]);
          if (release_v) umma_commit(&v_empty[slot]);
        }
        __syncwarp();
      };
      int idx = blockIdx.x;
      if (idx < n_items) {
        PairItem w = pair_item(sc, p.Skv, idx);
        uint32_t g = 0;
        int it = 0;
        mbar_wait(&q_full[0], 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        issue_qk(0, 0, 0, false);
        issue_qk(1, 0, 0, true);
        while (true) {
          const bool has_next = idx + n_ctas < n_items;
          for (int j = 0; j < w.n_blocks; ++j, ++g) {
            const uint32_t par = g & 1;
            const bool more = j + 1 < w.n_blocks;
            const bool next_qk = more || has_next;  // the block after this one: same item, or block 0 of the next
            const uint32_t ng = g + 1;
            const int nslot = ng % ATT2_KS;
            const int nqb = more ? (it & 1) : ((it + 1) & 1);
            if (next_qk) {
              if (!more) mbar_wait(&q_full[nqb], ((it + 1) >> 1) & 1);
              mbar_wait(&k_full[nslot], (ng / ATT2_KS) & 1);
              mbar_wait(&s_free[0], par);
              tc_fence_after();
              issue_qk(0, nqb, nslot, false);
            }
            if (j > 0) {
              mbar_wait(&p_full[1], par ^ 1);
              if (j == 1 && it > 0) mbar_wait(&o_free[1], (it - 1) & 1);
              tc_fence_after();
              issue_pv(1, (g - 1) % ATT2_KS, j != 1, true);
            }
            if (next_qk) {
              mbar_wait(&s_free[1], par);
              tc_fence_after();
              issue_qk(1, nqb, nslot, true);
            }
            mbar_wait(&v_full[g % ATT2_KS], (g / ATT2_KS) & 1);
            mbar_wait(&p_full[0], par);
            if (j == 0 && it > 0) mbar_wait(&o_free[0], (it - 1) & 1);
            tc_fence_after();
            issue_pv(0, g % ATT2_KS, j != 0, false);
            // O_A is final half a block before O_B: tile A's warpgroup goes through its epilogue and into the next
            // item while tile B still exponentiates, so the anti-phase of the two survives the item boundary
            if (!more) {
              if (elect_one()) umma_commit(&o_final[0]);
              __syncwarp();
            }
          }
          mbar_wait(&p_full[1], (g - 1) & 1);
          if (w.n_blocks == 1 && it > 0) mbar_wait(&o_free[1], (it - 1) & 1);
          tc_fence_after();
          issue_pv(1, (g - 1) % ATT2_KS, w.n_blocks != 1, true);
          if (elect_one()) {
            umma_commit(&o_final[1]);
            umma_commit(&q_empty[it & 1]);
          }
          __syncwarp();
          if (!has_next) break;
          idx += n_ctas;
          w = pair_item(sc, p.Skv, idx);
          ++it;
        }
      }
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
    uint32_t g = 0;
    int it = 0;
    // item-boundary stamps of CTA 0, per tile: trace[(t * 8 + it) * 8 + {0 item start, 1 first S ready, 2 block loop
    // done, 3 O final, 4 stored}]
    const bool tracer = TRACE && blockIdx.x == 0 && q == 0 && lane == 0;
    for (int idx = blockIdx.x; idx < n_items; idx += n_ctas, ++it) {
      const PairItem w = pair_item(sc, p.Skv, idx);
      float m_ref = -INFINITY;
      float l = 0.f;
      if (tracer && it < 8) trace[(t * 8 + it) * 8 + 0] = clock64();
      for (int j = 0; j < w.n_blocks; ++j, ++g) {
        mbar_wait(&s_full[t], g & 1);
        tc_fence_after();
        if (tracer && it < 8 && j == 0) trace[(t * 8 + it) * 8 + 1] = clock64();
        float x[ATT_BN];
        {
          uint32_t* xr = reinterpret_cast<uint32_t*>(x);
          tmem_ld_32x32b_x32(t_s + 0, xr + 0);
          tmem_ld_32x32b_x32(t_s + 32, xr + 32);
          tmem_ld_32x32b_x32(t_s + 64, xr + 64);
          tmem_ld_32x32b_x32(t_s + 96, xr + 96);
          // the wait for PV_t(j-1) sits before S_t is handed back on purpose (anti-phase of the two warpgroups; see
          // attn_pair_kernel)
          if (j > 0) mbar_wait(&pv_done[t], (g - 1) & 1);
          tmem_ld_wait();
          tc_fence_after();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        const int valid = p.Skv - (w.kb0 + j) * ATT_BN;
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
        if (j > 0) {
          if (__any_sync(0xffffffffu, need)) {
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
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const int i0 = ch * 16 + i;
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
          tmem_st_32x32b_x16(t_p + ch * 16, pk);
        }
        l = l * alpha + ((acc0.x + acc0.y) + (acc1.x + acc1.y));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }

      // ---- epilogue of the item
      if (tracer && it < 8) trace[(t * 8 + it) * 8 + 2] = clock64();
      mbar_wait(&o_final[t], it & 1);
      tc_fence_after();
      if (tracer && it < 8) trace[(t * 8 + it) * 8 + 3] = clock64();
      float o[ATT_D];
      {
        uint32_t* orr = reinterpret_cast<uint32_t*>(o);
        tmem_ld_32x32b_x32(t_o, orr);
        tmem_ld_32x32b_x32(t_o + 32, orr + 32);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);  // PV_t(0) of the next item may overwrite O_t
      const int row_in_unit = t * 128 + row_in_tile;
      bool store = true;
      if (w.nparts > 1) {
        const int su = w.unit - sc.n_full;
        const size_t slot = static_cast<size_t>(su) * w.nparts + w.part;
        float4* wo = reinterpret_cast<float4*>(sc.ws_o) + slot * (16 * 256) + row_in_unit;
#pragma unroll
        for (int c = 0; c < ATT_D / 4; ++c) wo[c * 256] = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
        reinterpret_cast<float2*>(sc.ws_ml)[slot * 256 + row_in_unit] = make_float2(m_ref, l);
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 128) {
          const unsigned int old = atomicAdd(sc.ws_cnt + su, 1u);
          const bool last = old == static_cast<unsigned int>(w.nparts - 1);
          if (last) sc.ws_cnt[su] = 0u;
          *last_flag = last ? 1u : 0u;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        store = *last_flag != 0u;
        asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone has read the flag before the next item rewrites it
        if (store) {
          __threadfence();
          const size_t slot0 = static_cast<size_t>(su) * w.nparts;
          const float2* mlp = reinterpret_cast<const float2*>(sc.ws_ml) + slot0 * 256 + row_in_unit;
          float2 ml[ATT_MAX_SPLIT];
          float m = -INFINITY;
#pragma unroll
          for (int pp = 0; pp < ATT_MAX_SPLIT; ++pp) {
            if (pp < w.nparts) {
              ml[pp] = __ldcg(mlp + pp * 256);
              m = fmaxf(m, ml[pp].x);
            }
          }
          l = 0.f;
#pragma unroll
          for (int c = 0; c < ATT_D; ++c) o[c] = 0.f;
#pragma unroll 1
          for (int pp = 0; pp < w.nparts; ++pp) {
            float wgt = 0.f;
#pragma unroll
            for (int k = 0; k < ATT_MAX_SPLIT; ++k)
              if (k == pp) wgt = ex2_approx(ml[k].x - m), l = fmaf(wgt, ml[k].y, l);
            const float4* src = reinterpret_cast<const float4*>(sc.ws_o) + (slot0 + pp) * (16 * 256) + row_in_unit;
            float4 f[ATT_D / 4];
#pragma unroll
            for (int c = 0; c < ATT_D / 4; ++c) f[c] = __ldcg(src + c * 256);
#pragma unroll
            for (int c = 0; c < ATT_D / 4; ++c) {
              o[4 * c] = fmaf(wgt, f[c].x, o[4 * c]);
              o[4 * c + 1] = fmaf(wgt, f[c].y, o[4 * c + 1]);
              o[4 * c + 2] = fmaf(wgt, f[c].z, o[4 * c + 2]);
              o[4 * c + 3] = fmaf(wgt, f[c].w, o[4 * c + 3]);
            }
          }
        }
      }
      // The warp's 32 rows x 64 bf16 go out as ONE TMA store from a swizzled staging slab. (Storing 16 bytes per lane
      // to 32 different rows cost ~2000 cycles of LSU back-pressure per item: profiles/r2_attn_trace.txt.) Rows >= S_q
      // are clipped by the tensor map.
      if (store) {
        const float inv_l = 1.f / l;
        uint8_t* slab = sO + (t * 4 + q) * 4096;
        if (lane == 0) tma_store_wait_read0();  // the previous item's store has finished reading the slab
        __syncwarp();
        uint8_t* rowp = slab + lane * 128;
#pragma unroll
        for (int c = 0; c < ATT_D; c += 8) {
          uint4 v;
          v.x = pack_bf16x2(o[c + 0] * inv_l, o[c + 1] * inv_l);
          v.y = pack_bf16x2(o[c + 2] * inv_l, o[c + 3] * inv_l);
          v.z = pack_bf16x2(o[c + 4] * inv_l, o[c + 5] * inv_l);
          v.w = pack_bf16x2(o[c + 6] * inv_l, o[c + 7] * inv_l);
          *reinterpret_cast<uint4*>(rowp + (((c >> 3) ^ (lane & 7)) << 4)) = v;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&mapO, slab, w.head * ATT_D, w.q_pair * 256 + t * 128 + q * 32, w.batch);
          tma_store_commit();
        }
      }
      if (tracer && it < 8) trace[(t * 8 + it) * 8 + 4] = clock64();
    }
    if (lane == 0) tma_store_wait_all0();
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

namespace {
// workspace layout of the split units of one launch (at most one wave of them): partial O | (m, l) | tickets
constexpr int64_t kWsSlots = 640;  // parts of all split units of one launch (<= 4 waves of short CTAs)
constexpr int64_t kWsOBytes = kWsSlots * 256 * mvd::ATT_D * 4;
constexpr int64_t kWsMlBytes = kWsSlots * 256 * 2 * 4;
constexpr int64_t kWsCntBytes = kWsSlots * 4;
}  // namespace

extern "C" int64_t mvd_attention_workspace_bytes(void) { return kWsOBytes + kWsMlBytes + kWsCntBytes; }

extern "C" int mvd_attention_bf16_ws(const void* q, int64_t ldq, int64_t q_batch_stride, const void* k, int64_t ldk,
                                     int64_t k_batch_stride, const void* v, int64_t ldv, int64_t v_batch_stride,
                                     void* out, int64_t ldo, int64_t o_batch_stride, int batch, int heads, int s_q,
                                     int s_kv, float scale, void* workspace, int64_t workspace_bytes, int co_units,
                                     void* stream) {
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
  MVD_CHECK(workspace == nullptr || ((reinterpret_cast<uintptr_t>(workspace) & 15) == 0 &&
                                     workspace_bytes >= mvd_attention_workspace_bytes()),
            "attention: workspace must be 16-byte aligned and hold mvd_attention_workspace_bytes() = %lld bytes",
            static_cast<long long>(mvd_attention_workspace_bytes()));

  CUtensorMap mQ, mK, mV;
  // K/V may be shared by all batch entries (batch stride 0: the cross-view reference K/V of configs[3]); a zero
  // stride is not encodable in a tensor map, so the batch dimension is then dropped to extent 1 and every CTA reads
  // coordinate 0.
  auto mk = [&](CUtensorMap* m, const void* ptr, int64_t ld, int64_t bstride, int S) -> int {
    const bool bcast = bstride == 0;
    const uint64_t dims[3] = {static_cast<uint64_t>(heads) * ATT_D, static_cast<uint64_t>(S),
                              static_cast<uint64_t>(bcast ? 1 : batch)};
    const uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2,
                                 static_cast<uint64_t>(bcast ? static_cast<int64_t>(S) * ld : bstride) * 2};
    const uint32_t box[3] = {ATT_D, 128, 1};
    return make_tmap_bf16(m, ptr, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  MVD_CHECK(q_batch_stride != 0 && o_batch_stride != 0, "attention: q / out need a batch stride");
  MVD_CHECK((k_batch_stride == 0) == (v_batch_stride == 0), "attention: k and v must both be shared or both per batch");
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
  a.kv_batch_mul = k_batch_stride == 0 ? 0 : 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sms = sm_count();

  // long query sequences: two Q tiles per CTA (two softmax warpgroups); short ones: one tile per CTA so that small
  // sites still spread over the SMs
  const long q_pairs = (s_q + 255) / 256;
  const long units = q_pairs * heads * batch;
  // with a workspace even a launch of fewer units than SMs fills the machine (every unit is split along S_kv)
  const bool two_tiles = (s_q >= 512) && units >= (workspace ? static_cast<long>(sms) / 4 : 2L * sms);
  if (two_tiles) {
    MVD_CHECK(units < (1L << 30), "attention: too many tiles");
    PairSched sc;
    memset(&sc, 0, sizeof(sc));
    sc.q_pairs = static_cast<int>(q_pairs);
    sc.heads = heads;
    sc.n_full = static_cast<int>(units);
    sc.split = 1;
    const int n_blocks = (s_kv + ATT_BN - 1) / ATT_BN;
    // co_units: 256-row units of OTHER launches the caller runs concurrently (the adapter's second attention branch on
    // a side stream): the machine's last wave is then shared, and only what this launch contributes to it is split
    MVD_CHECK(co_units >= 0, "attention: co_units must be >= 0");
    const int rem = static_cast<int>((units + co_units) % sms) <= units ? static_cast<int>((units + co_units) % sms)
                                                                         : static_cast<int>(units);
    static const bool split_on = [] {
      const char* e = getenv("MVD_ATTN_SPLIT");
      return e == nullptr || e[0] != '0';
    }();
    if (workspace && split_on && rem > 0) {
      // `tail` units share the last wave. Split them into s parts each so that the wave is made of short CTAs:
      // cost model in KV blocks, with kFixed blocks of fixed cost per CTA (prologue, partial write, merge):
      //   waves(s) * (n_blocks / s + kFixed),  waves(s) = ceil(tail * s / SMs)   (one wave when the tail is < 1 wave of units)
      // This also covers launches with fewer units than SMs (view-sharded ranks, B = 1 or 2): tail = all units.
      const int tail = rem;
      int best = 1;
      constexpr double kFixed = 4.0;  // prologue + partial write + merge of one CTA, in KV blocks (measured ~8 us)
      double best_cost = (static_cast<double>((tail + sms - 1) / sms)) * (n_blocks + kFixed);
      for (int sp = 2; sp <= ATT_MAX_SPLIT; ++sp) {
        if (n_blocks / sp < 2 || static_cast<int64_t>(tail) * sp > kWsSlots) break;
        const double waves = static_cast<double>((static_cast<int64_t>(tail) * sp + sms - 1) / sms);
        const double cost = waves * (static_cast<double>(n_blocks) / sp + kFixed);
        if (cost < best_cost * 0.9) {
          best_cost = cost;
          best = sp;
        }
      }
      if (best >= 2) {
        sc.n_full = static_cast<int>(units) - tail;
        sc.split = best;
        char* w = static_cast<char*>(workspace);
        sc.ws_o = reinterpret_cast<float*>(w);
        sc.ws_ml = reinterpret_cast<float*>(w + kWsOBytes);
        sc.ws_cnt = reinterpret_cast<unsigned int*>(w + kWsOBytes + kWsMlBytes);
      }
    }
    const dim3 grid(sc.n_full + (static_cast<int>(units) - sc.n_full) * sc.split);
    const char* tp = getenv("MVD_ATTN_TRACE_PTR");
    long long* trace = reinterpret_cast<long long*>(tp ? strtoull(tp, nullptr, 0) : 0ull);
    static const int poly = [] {
      const char* e = getenv("MVD_ATTN_POLY8");
      return e ? atoi(e) : 0;
    }();
    // at least one whole wave of units: persistent CTAs (one per SM) walk the items
    static const bool persist_on = [] {
      const char* e = getenv("MVD_ATTN_PERSIST");
      return e == nullptr || e[0] != '0';
    }();
    // (long items amortise the per-item costs by themselves, and measured slower in this form: configs[3]'s 576-block
    // reference attention 2.20 vs 2.09 ms, step 125.5 vs 122.1 ms)
    if (persist_on && units >= sms && n_blocks <= ATT_PERSIST_MAX_BLOCKS) {
      const int n_items = static_cast<int>(grid.x);
      const dim3 pgrid(n_items < sms ? n_items : sms);
      CUtensorMap mO;  // output rows, one warp slab (32 rows x 64 columns) per store
      {
        const uint64_t dims[3] = {static_cast<uint64_t>(heads) * ATT_D, static_cast<uint64_t>(s_q),
                                  static_cast<uint64_t>(batch)};
        const uint64_t strides[2] = {static_cast<uint64_t>(ldo) * 2, static_cast<uint64_t>(o_batch_stride) * 2};
        const uint32_t box[3] = {ATT_D, 32, 1};
        if (int e = make_tmap_bf16(&mO, out, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
      }
      if (trace) {
        MVD_CUDA(cudaFuncSetAttribute(attn_pair_persist_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      ATT3_SMEM));
        MVD_CUDA(launch_pdl(attn_pair_persist_kernel<0, true>, pgrid, dim3(ATT2_THREADS), ATT3_SMEM, st, mQ, mK, mV, mO,
                            a, sc, n_items, trace));
        MVD_CUDA(cudaGetLastError());
        count_launches(1);
        return MVD_OK;
      }
#define MVD_PAIR_P(POLY)                                                                                                \
  do {                                                                                                                  \
    MVD_CUDA(cudaFuncSetAttribute(attn_pair_persist_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                  ATT3_SMEM));                                                                          \
    MVD_CUDA(launch_pdl(attn_pair_persist_kernel<POLY>, pgrid, dim3(ATT2_THREADS), ATT3_SMEM, st, mQ, mK, mV, mO, a,    \
                        sc, n_items, static_cast<long long*>(nullptr)));                                                                                      \
  } while (0)
      if (poly == 1) MVD_PAIR_P(1);
      else if (poly == 2) MVD_PAIR_P(2);
      else MVD_PAIR_P(0);
#undef MVD_PAIR_P
      MVD_CUDA(cudaGetLastError());
      count_launches(1);
      return MVD_OK;
    }
#define MVD_PAIR(POLY, TR)                                                                                              \
  do {                                                                                                                  \
    MVD_CUDA(cudaFuncSetAttribute(attn_pair_kernel<POLY, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM)); \
    MVD_CUDA(launch_pdl(attn_pair_kernel<POLY, TR>, grid, dim3(ATT2_THREADS), ATT2_SMEM, st, mQ, mK, mV, a, sc, trace)); \
  } while (0)
    if (trace) MVD_PAIR(0, true);
    else if (poly == 1) MVD_PAIR(1, false);
    else if (poly == 2) MVD_PAIR(2, false);
    else MVD_PAIR(0, false);
#undef MVD_PAIR
  } else {
    MVD_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    dim3 grid((s_q + ATT_BM - 1) / ATT_BM, heads, batch);
    MVD_CUDA(launch_pdl(attn_fwd_kernel, grid, dim3(192), ATT_SMEM, st, mQ, mK, mV, a));
  }
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

extern "C" int mvd_attention_bf16(const void* q, int64_t ldq, int64_t q_batch_stride, const void* k, int64_t ldk,
                                  int64_t k_batch_stride, const void* v, int64_t ldv, int64_t v_batch_stride,
                                  void* out, int64_t ldo, int64_t o_batch_stride, int batch, int heads, int s_q,
                                  int s_kv, float scale, void* stream) {
  return mvd_attention_bf16_ws(q, ldq, q_batch_stride, k, ldk, k_batch_stride, v, ldv, v_batch_stride, out, ldo,
                               o_batch_stride, batch, heads, s_q, s_kv, scale, nullptr, 0, 0, stream);
}
