// gemm.cu — persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
// One kernel serves every dense contraction of the denoise step that is not attention:
//   * Linear layers (diffusers Attention.to_q/k/v/out, Transformer2DModel.proj_in/out, FeedForward;
//     reference adapter projections src/models/attention.py:125-132,157): the 1-tap case.
//   * 3x3 convolutions of ResnetBlock2D / Upsample2D / conv shortcut: 9 taps,
//     the A operand is fetched by 4-D TMA boxes over the NHWC activation, shifted per tap; TMA
//     out-of-bounds zero fill implements the padding=1 halo.
//   * Stride-2 3x3 convolutions of Downsample2D: the NHWC input is viewed as [N, H/2, 2, W/2, 2C] (5-D)
//     so every tap is again one dense TMA box.
//
// Layout: activations are NHWC bf16 ([rows, C] row-major == K-major A operand), weights are [N, K]
// row-major bf16 (K-major B operand; for convs K = tap*Cin + c). Accumulation fp32 in TMEM.
//
// CTA = 320 threads: warp 0 TMA producer, warp 1 TMEM owner + single-thread MMA issuer, warps 2..5 and 6..9 two
// epilogue warpgroups (TMEM -> registers -> fused bias / per-image bias / residual / GEGLU -> bf16 -> smem -> TMA
// store), one per TMEM accumulator stage. Persistent: grid = min(#tiles, #SMs); the epilogues of tiles i and i+1
// overlap each other and the main loop of tile i+2's predecessor stage.
#include "tc.cuh"
#include "host_common.h"
#include "../../include/mvd_b200.h"

namespace mvd {

constexpr int BM = 128;  // tile rows (output pixels)
constexpr int BK = 64;   // bf16 elements per k-block = one 128-B swizzle row
constexpr int EPI_BUFS = 4;
constexpr int EPI_BUF_BYTES = 32 * 32 * 2;  // 32 rows x 32 bf16
// Two epilogue warpgroups, one per TMEM accumulator stage (warpgroup w drains the CTA's tiles w, w+2, ...): with a
// single warpgroup the small-K GEMMs (K = 320 projections, GEGLU) were bound by the epilogue's instruction issue
// (one warp per scheduler), not by the MMA pipe.
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 32 * (2 + EPI_WARPS);

struct GemmArgs {
  int tiles_total, tiles_n;
  int tiles_x, tiles_y;  // m-tile -> (x, y, img) with x fastest
  int TW, TH, TN;        // tile = TN images x TH rows x TW pixels (=128)
  int bx, by;            // per-epilogue-warp store box (bx * by * bn = 32 pixels)
  int ntaps;             // 1 (linear / 1x1) or 9 (3x3)
  int kc_per_tap;        // k-blocks per tap (Cin / 64, both sources)
  int kc_a1;             // k-blocks per tap that come from source 1 (rest from source 2)
  int cin;               // Cin: weight K offset of tap t is t * cin
  int c_s2;              // stride-2 mode: channel count C of the 5-D view (inner dim is 2C)
  int N;                 // output columns
  int geglu;             // 1: tile columns are [a | g] halves, out = a * gelu(g), N/2 output columns
  int has_res;
  int rows_per_img;      // img index for img_bias: n_coord + x_coord / rows_per_img
  int img_bias_ld;
  int img_max;           // clamp for the img index (padded rows of the last tile)
  const __nv_bfloat16* bias;  // [N] (permuted like the weight rows when geglu)
  const float* img_bias;      // [imgs, img_bias_ld] fp32 (e.g. time-embedding projection)
  // --- fused neighbours of the GEMM (all optional)
  // LayerNorm of the A rows folded into the GEMM: the weights carry gamma, and per output row
  //   v = rstd * (acc - mean * ln_colsum[col]);  mean / rstd from ln_parts partial (sum, sum of squares) pairs
  const float2* ln_stats;     // [rows][ln_parts]
  const float* ln_colsum;     // [N] fp32: sum_k of the (gamma-scaled, bf16) weight row
  int ln_parts;
  float ln_inv_k, ln_eps;
  // row statistics of the bf16 OUTPUT for the LayerNorm that follows: partial (sum, sum of squares) per row and
  // column tile, [rows][tiles_n]
  float2* stats_out;
  int rows_total;             // M (linear mode)
  // FiLM on the output (camera modulation of a block output): v = v * film_scale[img][col] + film_shift[img][col]
  const float* film_scale;
  const float* film_shift;
  int film_ld;
  // stream-K (SPLIT kernels): tiles [0, sk_first) are whole-tile work items; the k-blocks of the remaining tiles — the
  // last, partial wave, or all tiles of a launch smaller than the machine — are shared evenly by CTAs [0, sk_ctas)
  int sk_first;
  int sk_ctas;
  int sk_units;              // (tiles_total - sk_first) * k_blocks
  int sk_slots;              // partial-tile slots per stream-K tile (>= the CTAs that can touch one tile)
  float4* ws_partial;        // [tile - sk_first][slot][BN/4][128 rows] fp32 partial accumulators (lane = row: coalesced)
  unsigned int* ws_tickets;  // [tile - sk_first][4] arrival counters, one per 32-row warp slab; zero between launches
};

// One unit of a CTA's work: k-blocks [kb0, kb1) of an output tile. nsl == 1: the whole tile (normal epilogue);
// nsl > 1: segment `slice` (in k order) of the nsl segments that different CTAs contribute to the tile.
struct WorkItem {
  int tile, kb0, kb1, slice, nsl;
};

// The sequence of work items of CTA `unit`, identical in the producer, MMA and epilogue roles. Stream-K segments come
// first (their partial exchange then overlaps the whole-tile work of the other CTAs), whole tiles round-robin after.
// CTA i of the stream-K set owns units [floor(i T / G), floor((i + 1) T / G)) of the T = tiles * k_blocks k-block units;
// the CTA that owns unit u is floor(((u + 1) G - 1) / T)  (host guarantees T >= G and T * G < 2^31).
template <bool SPLIT>
struct WorkIter {
  int u0, u1, t, step, n_direct, k_blocks, unit;
  __device__ __forceinline__ WorkIter(const GemmArgs& p, int k_blocks_, int unit_, int n_units)
      : u0(0), u1(0), t(unit_), step(n_units), n_direct(SPLIT ? p.sk_first : p.tiles_total), k_blocks(k_blocks_),
        unit(unit_) {
    if (SPLIT && unit_ < p.sk_ctas) {
      u0 = unit_ * p.sk_units / p.sk_ctas;
      u1 = (unit_ + 1) * p.sk_units / p.sk_ctas;
    }
  }
  __device__ __forceinline__ bool next(const GemmArgs& p, WorkItem& w) {
    if (SPLIT && u0 < u1) {
      const int rel = u0 / k_blocks;
      const int base = rel * k_blocks;
      w.tile = p.sk_first + rel;
      w.kb0 = u0 - base;
      const int len = min(k_blocks - w.kb0, u1 - u0);
      w.kb1 = w.kb0 + len;
      const int c_first = ((base + 1) * p.sk_ctas - 1) / p.sk_units;
      const int c_last = ((base + k_blocks) * p.sk_ctas - 1) / p.sk_units;
      w.slice = unit - c_first;
      w.nsl = c_last - c_first + 1;
      u0 += len;
      return true;
    }
    if (t < n_direct) {
      w.tile = t;
      w.kb0 = 0;
      w.kb1 = k_blocks;
      w.slice = 0;
      w.nsl = 1;
      t += step;
      return true;
    }
    return false;
  }
};

constexpr int WS_MAX_KBLOCKS = 5;  // weight-stationary tiles: K <= 320

// WS (weight-stationary, K <= 320): the CTA's weight tile [BN x K] stays in shared memory for the whole launch (every
// CTA keeps ONE column tile: the grid is a multiple of the column-tile count) and only A streams through the ring.
template <int BN, bool WS = false>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int B_RES_BYTES = WS ? WS_MAX_KBLOCKS * B_BYTES : 0;
  static constexpr int STAGE_BYTES = WS ? A_BYTES : A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = EPI_WARPS * EPI_BUFS * EPI_BUF_BYTES;  // 64 KB
  static constexpr int STAGES = WS ? 5 : (BN <= 64) ? 6 : (BN <= 128) ? 5 : (BN <= 160) ? 4 : 3;
  static_assert(STAGE_BYTES % 1024 == 0, "128B-swizzled tiles must stay 1024-byte aligned");
  static constexpr int BAR_BYTES = 1024;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + B_RES_BYTES + EPI_BYTES + BAR_BYTES + 1024 /*align slack*/;
  static constexpr int ACC_STRIDE = (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static_assert(TOTAL <= 227 * 1024, "shared memory budget of one CTA");
};

// Exact (erf) GELU, x * Phi(x), as diffusers' GEGLU uses. The GEGLU epilogue is instruction-bound (4 epilogue warps,
// 128 activations per thread per tile), and libm's erff costs ~30 instructions; this is Abramowitz-Stegun 7.1.26,
// erfc(z) = t*(a1 + t*(a2 + ...))*exp(-z^2), t = 1/(1 + p*z), |error| < 1.5e-7 (4.2e-7 on GELU in fp32 including the
// approximate rcp/ex2, i.e. four orders of magnitude below the bf16 rounding of the result), ~17 instructions.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  const float half_erfc = 0.5f * (p * t) * e;  // 0.5 * erfc(|x| / sqrt 2) = Phi(-|x|)
  return x * (x >= 0.f ? 1.0f - half_erfc : half_erfc);
}

// Round 2 measured and removed CTA-pair tiles (cta_group::2, 256 x BN; 5-20 % slower at every M = 32768 linear, step
// 13.74 vs 13.26 ms; profiles/r2_gemm_variants.txt).
// SPLIT: stream-K for the part of a launch that does not fill the machine — all tiles of the 16x16 / 8x8 levels and of
// a view-sharded rank (1-2 samples: 20-120 tiles with 45-360 k-blocks each), and the last partial wave of the big
// launches. The k-blocks of those tiles are shared evenly by the CTAs (WorkIter): a CTA's share covers the end of one
// tile and the beginning of the next. Each segment's fp32 partial tile goes to a workspace (column-chunk major, so a
// warp's 32 rows are contiguous), and per 32-row slab the LAST arriving warp (ticket counter, re-armed for the next
// launch) adds the segments in k order — deterministic — and runs the normal epilogue on the sum; a tile that one CTA
// covers alone takes the normal path. Round 1's split-K (row-major partials at M = 512, short k-loops) measured
// slower and was removed; uniform k-slices per tile (first half of round 2) could not use 148 SMs for 80 tiles.
template <int BN, bool S2, bool WS = false, bool SPLIT = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_conv_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA2,
                 const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapOut,
                 const __grid_constant__ CUtensorMap mapRes, const GemmArgs p) {
  using L = SmemLayout<BN, WS>;
  static_assert(!WS || (!S2 && !SPLIT), "weight-stationary tiles: 1-tap only, un-split");
  const int unit = blockIdx.x;
  const int n_units = gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  [[maybe_unused]] uint8_t* b_res = smem + L::STAGES * L::STAGE_BYTES;  // WS: resident weight tile, k-block major
  uint8_t* epi_smem = smem + L::STAGES * L::STAGE_BYTES + L::B_RES_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + L::EPI_BYTES);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = bars + L::STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * L::STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint64_t* res_bar = tmem_empty + 2;          // [EPI_WARPS][EPI_BUFS]
  uint64_t* b_full = res_bar + EPI_WARPS * EPI_BUFS;  // WS: the resident weight tile has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform role index
  const int lane = threadIdx.x & 31;
  const int k_blocks = p.ntaps * p.kc_per_tap;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    tma_prefetch_desc(&mapOut);
    for (int s = 0; s < L::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    for (int i = 0; i < EPI_WARPS * EPI_BUFS; ++i) mbar_init(&res_bar[i], 1);
    if constexpr (WS) mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, L::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; from here on we read its results
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if constexpr (WS) {
        // gridDim.x is a multiple of tiles_n, so t % tiles_n is the same for every tile of this CTA
        const int n_tile_fixed = static_cast<int>(blockIdx.x) % p.tiles_n;
        mbar_arrive_expect_tx(b_full, static_cast<uint32_t>(k_blocks) * L::B_BYTES);
        for (int kb = 0; kb < k_blocks; ++kb)
          tma_load_2d(b_res + kb * L::B_BYTES, &mapB, b_full, kb * BK, n_tile_fixed * BN);
      }
      WorkIter<SPLIT> work(p, k_blocks, unit, n_units);
      WorkItem wk;
      while (work.next(p, wk)) {
        const int tile = wk.tile, kb0 = wk.kb0, kb1 = wk.kb1;
        const int n_tile = tile % p.tiles_n;
        int m_tile = tile / p.tiles_n;
        const int x0 = (m_tile % p.tiles_x) * p.TW;
        m_tile /= p.tiles_x;
        const int y0 = (m_tile % p.tiles_y) * p.TH;
        const int n0 = (m_tile / p.tiles_y) * p.TN;
        int tap = kb0 / p.kc_per_tap, kc = kb0 - tap * p.kc_per_tap;
        for (int kb = kb0; kb < kb1; ++kb) {
          const int ky = (p.ntaps == 9) ? tap / 3 : 1;
          const int kx = (p.ntaps == 9) ? tap % 3 : 1;
          {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::STAGE_BYTES;
            uint8_t* sb = sa + L::A_BYTES;
            if constexpr (WS) {
              mbar_arrive_expect_tx(&full[stage], L::A_BYTES);
              tma_load_4d(sa, &mapA, &full[stage], kc * BK, x0 + kx - 1, y0 + ky - 1, n0);
            } else {
              mbar_arrive_expect_tx(&full[stage], L::STAGE_BYTES);
              if constexpr (S2) {
                // input pixel (2*oy + ky - 1, 2*ox + kx - 1) in the [N, H/2, 2, W/2, 2C] view
                const int px = (kx == 1) ? 0 : 1, py = (ky == 1) ? 0 : 1;
                const int wx = x0 + ((kx == 0) ? -1 : 0), hy = y0 + ((ky == 0) ? -1 : 0);
                tma_load_5d(sa, &mapA, &full[stage], px * p.c_s2 + kc * BK, wx, py, hy, n0);
              } else {
                if (kc < p.kc_a1)
                  tma_load_4d(sa, &mapA, &full[stage], kc * BK, x0 + kx - 1, y0 + ky - 1, n0);
                else
                  tma_load_4d(sa, &mapA2, &full[stage], (kc - p.kc_a1) * BK, x0 + kx - 1, y0 + ky - 1, n0);
              }
              tma_load_2d(sb, &mapB, &full[stage], tap * p.cin + kc * BK, n_tile * BN);
            }
            if (++stage == L::STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (++kc == p.kc_per_tap) {
            kc = 0;
            ++tap;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Warp-uniform control flow: all lanes wait on the barriers and build the descriptors (uniform registers),
    // one elected lane issues the tcgen05.mma / commit instructions.
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem_base = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if constexpr (WS) {
      mbar_wait(b_full, 0);
      tc_fence_after();
    }
    WorkIter<SPLIT> work(p, k_blocks, unit, n_units);
    WorkItem wk;
    for (; work.next(p, wk); ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tb + acc * L::ACC_STRIDE;
      const int kb0 = wk.kb0, kb1 = wk.kb1;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
        const uint64_t adesc = umma_desc_sw128(sa);
        const uint64_t bdesc = WS ? umma_desc_sw128(smem_base + L::STAGES * L::STAGE_BYTES + kb * L::B_BYTES)
                                  : umma_desc_sw128(sa + L::A_BYTES);
        if (elect_one()) {
          // +32 B per K=16 step inside the 128-B swizzle row (start-address field is addr >> 4)
          umma_ss(d_tmem, adesc, bdesc, idesc, kb != kb0);
#pragma unroll
          for (int k = 1; k < BK / 16; ++k) umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == L::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);
      __syncwarp();
    }
  } else if (warp >= 2) {
    // ===================== epilogue warps =====================
    const int q = warp & 3;          // TMEM lane quadrant this warp may access
    const int wg = (warp - 2) >> 2;  // epilogue warpgroup = the accumulator stage it drains
    uint8_t* my_smem = epi_smem + (wg * 4 + q) * (EPI_BUFS * EPI_BUF_BYTES);
    uint64_t* my_res_bar = res_bar + (wg * 4 + q) * EPI_BUFS;
    const int r0 = q * 32;  // first tile row of this warp
    // position of this warp's 32-pixel slab inside the tile
    const int wx_off = r0 % p.TW;
    const int wy_off = (r0 / p.TW) % p.TH;
    const int wn_off = r0 / (p.TW * p.TH);
    // position of this lane's pixel inside the tile (for the per-image bias)
    const int row = r0 + lane;
    const int lx = row % p.TW;
    const int ln = row / (p.TW * p.TH);
    const int n_chunks = p.geglu ? BN / 64 : BN / 32;
    const int out_cols_per_tile = p.geglu ? BN / 2 : BN;
    const int n_out = p.geglu ? p.N / 2 : p.N;

    uint32_t g = 0;  // running chunk counter -> staging buffer + barrier parity
    int it = 0;      // index of the work item in this CTA's sequence: this warpgroup drains stage it & 1 == wg
    WorkIter<SPLIT> work(p, k_blocks, unit, n_units);
    WorkItem wk;
    for (; work.next(p, wk); ++it) {
      if ((it & 1) != wg) continue;
      const int tile = wk.tile;
      const int n_tile = tile % p.tiles_n;
      int m_tile = tile / p.tiles_n;
      const int x0 = (m_tile % p.tiles_x) * p.TW;
      m_tile /= p.tiles_x;
      const int y0 = (m_tile % p.tiles_y) * p.TH;
      const int n0 = (m_tile / p.tiles_y) * p.TN;
      const int sx = x0 + wx_off, sy = y0 + wy_off, sn = n0 + wn_off;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int col_base = n_tile * out_cols_per_tile;

      [[maybe_unused]] const float4* part_rd = nullptr;
      if (SPLIT && wk.nsl > 1) {
        // publish this segment's fp32 partial tile, hand the accumulator back, take a ticket for the 32-row slab
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_part = tmem_base + acc * L::ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16);
        if (wk.nsl > p.sk_slots) __trap();  // host / device disagree on the segment count: never corrupt silently
        const size_t slot0 = static_cast<size_t>(tile - p.sk_first) * p.sk_slots;
        float4* part = p.ws_partial + (slot0 + wk.slice) * (BN / 4 * BM) + row;
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_part + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            part[static_cast<size_t>(c * 8 + j) * BM] =
                make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                            __uint_as_float(r[4 * j + 3]));
        }
        tc_fence_before();
        __threadfence();  // partial rows visible device-wide before the ticket is taken
        __syncwarp();
        int last = 0;
        if (lane == 0) {
          mbar_arrive(&tmem_empty[acc]);
          unsigned int* tk = p.ws_tickets + (tile - p.sk_first) * 4 + q;
          last = atomicAdd(tk, 1u) == static_cast<unsigned int>(wk.nsl - 1);
          if (last) *tk = 0u;  // re-arm for the next launch
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (!last) continue;
        __threadfence();
        part_rd = p.ws_partial + slot0 * (BN / 4 * BM) + row;
      }

      // prefetch residual slabs for the first two chunks (their buffers are free: at most one store
      // group from the previous tile may still be reading, and it is neither of these two buffers
      // after the wait below)
      if (p.has_res) {
        if (lane == 0) {
          tma_store_wait_read0();
          for (int c = 0; c < 2 && c < n_chunks; ++c) {
            const uint32_t gb = (g + c) % EPI_BUFS;
            mbar_arrive_expect_tx(&my_res_bar[gb], EPI_BUF_BYTES);
            tma_load_4d(my_smem + gb * EPI_BUF_BYTES, &mapRes, &my_res_bar[gb], col_base + c * 32, sx, sy, sn);
          }
        }
        __syncwarp();
      }

      if (part_rd == nullptr) {
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
      }
      const uint32_t t_acc = tmem_base + acc * L::ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16);
      int img = ln + n0 + (x0 + lx) / p.rows_per_img;
      img = img < p.img_max ? img : p.img_max;
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (p.ln_stats != nullptr) {  // linear mode: this lane's row is x0 + lx
        const int grow = (x0 + lx) < p.rows_total ? (x0 + lx) : p.rows_total - 1;
        const float2* st = p.ln_stats + static_cast<size_t>(grow) * p.ln_parts;
        float s1 = 0.f, s2 = 0.f;
        for (int pp = 0; pp < p.ln_parts; ++pp) {  // fixed order: deterministic
          const float2 t2 = __ldg(st + pp);
          s1 += t2.x;
          s2 += t2.y;
        }
        ln_mean = s1 * p.ln_inv_k;
        ln_rstd = rsqrtf(fmaxf(fmaf(s2, p.ln_inv_k, -ln_mean * ln_mean), 0.f) + p.ln_eps);
      }
      const float ln_k2 = -ln_rstd * ln_mean;  // v = rstd * acc + (k2 * colsum + c): two FFMAs per element
      float st_sum = 0.f, st_sq = 0.f;

      for (int c = 0; c < n_chunks; ++c, ++g) {
        const uint32_t buf = g % EPI_BUFS;
        const uint32_t buf_parity = (g / EPI_BUFS) & 1;
        uint8_t* sbuf = my_smem + buf * EPI_BUF_BYTES;
        const int col0 = col_base + c * 32;  // first output column of this chunk

        // buffer (g+2)%4 was last stored by chunk g-2: allow only the newest store group to be pending
        if (lane == 0) {
          // has_res: buffer (g+2)%4 (last stored by chunk g-2) is about to receive a residual slab -> at most the
          // newest store may be pending; otherwise only buffer g%4 (chunk g-4) must be drained -> three may be.
          if (p.has_res)
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else
            asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
          if (p.has_res && c + 2 < n_chunks) {
            const uint32_t gb = (g + 2) % EPI_BUFS;
            mbar_arrive_expect_tx(&my_res_bar[gb], EPI_BUF_BYTES);
            tma_load_4d(my_smem + gb * EPI_BUF_BYTES, &mapRes, &my_res_bar[gb], col0 + 64, sx, sy, sn);
          }
        }
        __syncwarp();

        float v[32];
        if (!p.geglu) {
          if (SPLIT && part_rd != nullptr) {
            // segments in k order (fixed summation order whichever CTA arrived last); L2-coherent, coalesced loads
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
            for (int sl = 0; sl < wk.nsl; ++sl) {
              const float4* src = part_rd + static_cast<size_t>(sl) * (BN / 4 * BM) + static_cast<size_t>(c * 8) * BM;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 f = __ldcg(src + static_cast<size_t>(j) * BM);
                v[4 * j] += f.x; v[4 * j + 1] += f.y; v[4 * j + 2] += f.z; v[4 * j + 3] += f.w;
              }
            }
          } else {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_acc + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          }
          if (col0 + 32 <= p.N) {  // N is a multiple of 32: a chunk is either fully inside or fully outside
            if (p.ln_stats != nullptr) {
              // LayerNorm fold fused with the fp32 constant (W.beta + bias, passed as the row-group bias row)
              const float4* cs = reinterpret_cast<const float4*>(p.ln_colsum + col0);
              const float4* cb = reinterpret_cast<const float4*>(p.img_bias + static_cast<size_t>(img) * p.img_bias_ld + col0);
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) {
                const float4 f = __ldg(cs + q4), c4 = __ldg(cb + q4);
                v[q4 * 4 + 0] = fmaf(v[q4 * 4 + 0], ln_rstd, fmaf(f.x, ln_k2, c4.x));
                v[q4 * 4 + 1] = fmaf(v[q4 * 4 + 1], ln_rstd, fmaf(f.y, ln_k2, c4.y));
                v[q4 * 4 + 2] = fmaf(v[q4 * 4 + 2], ln_rstd, fmaf(f.z, ln_k2, c4.z));
                v[q4 * 4 + 3] = fmaf(v[q4 * 4 + 3], ln_rstd, fmaf(f.w, ln_k2, c4.w));
              }
            } else {
            if (p.bias != nullptr) {
              const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col0);  // 64-byte aligned
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const uint4 bv = __ldg(bp + q4);
                const float2 b0 = unpack_bf16x2(bv.x), b1 = unpack_bf16x2(bv.y), b2 = unpack_bf16x2(bv.z),
                             b3 = unpack_bf16x2(bv.w);
                v[q4 * 8 + 0] += b0.x; v[q4 * 8 + 1] += b0.y; v[q4 * 8 + 2] += b1.x; v[q4 * 8 + 3] += b1.y;
                v[q4 * 8 + 4] += b2.x; v[q4 * 8 + 5] += b2.y; v[q4 * 8 + 6] += b3.x; v[q4 * 8 + 7] += b3.y;
              }
            }
            if (p.img_bias != nullptr) {
              const float4* ib = reinterpret_cast<const float4*>(p.img_bias + static_cast<size_t>(img) * p.img_bias_ld +
                                                                 col0);
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) {
                const float4 f = __ldg(ib + q4);
                v[q4 * 4 + 0] += f.x; v[q4 * 4 + 1] += f.y; v[q4 * 4 + 2] += f.z; v[q4 * 4 + 3] += f.w;
              }
            }
            }
          }
        } else {
          uint32_t ra[32], rg[32];
          tmem_ld_32x32b_x32(t_acc + c * 32, ra);
          tmem_ld_32x32b_x32(t_acc + BN / 2 + c * 32, rg);
          tmem_ld_wait();
          const int wa = n_tile * BN + c * 32;  // permuted weight-row index of the 'a' columns
          const int wg = wa + BN / 2;
          float ba[32], bg[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) ba[j] = bg[j] = 0.f;
          if (p.bias != nullptr) {
            const uint4* pa = reinterpret_cast<const uint4*>(p.bias + wa);
            const uint4* pg = reinterpret_cast<const uint4*>(p.bias + wg);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const uint4 av = __ldg(pa + q4), gv = __ldg(pg + q4);
              const float2 a0 = unpack_bf16x2(av.x), a1 = unpack_bf16x2(av.y), a2 = unpack_bf16x2(av.z),
                           a3 = unpack_bf16x2(av.w);
              const float2 g0 = unpack_bf16x2(gv.x), g1 = unpack_bf16x2(gv.y), g2 = unpack_bf16x2(gv.z),
                           g3 = unpack_bf16x2(gv.w);
              ba[q4 * 8 + 0] = a0.x; ba[q4 * 8 + 1] = a0.y; ba[q4 * 8 + 2] = a1.x; ba[q4 * 8 + 3] = a1.y;
              ba[q4 * 8 + 4] = a2.x; ba[q4 * 8 + 5] = a2.y; ba[q4 * 8 + 6] = a3.x; ba[q4 * 8 + 7] = a3.y;
              bg[q4 * 8 + 0] = g0.x; bg[q4 * 8 + 1] = g0.y; bg[q4 * 8 + 2] = g1.x; bg[q4 * 8 + 3] = g1.y;
              bg[q4 * 8 + 4] = g2.x; bg[q4 * 8 + 5] = g2.y; bg[q4 * 8 + 6] = g3.x; bg[q4 * 8 + 7] = g3.y;
            }
          }
          if (p.ln_stats != nullptr) {  // LayerNorm folded into both halves (column sums follow the weight rows)
            const float4* ca = reinterpret_cast<const float4*>(p.ln_colsum + wa);
            const float4* cg = reinterpret_cast<const float4*>(p.ln_colsum + wg);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              const float4 fa = __ldg(ca + q4), fg = __ldg(cg + q4);
              ba[q4 * 4 + 0] = fmaf(fa.x, ln_k2, ba[q4 * 4 + 0]); bg[q4 * 4 + 0] = fmaf(fg.x, ln_k2, bg[q4 * 4 + 0]);
              ba[q4 * 4 + 1] = fmaf(fa.y, ln_k2, ba[q4 * 4 + 1]); bg[q4 * 4 + 1] = fmaf(fg.y, ln_k2, bg[q4 * 4 + 1]);
              ba[q4 * 4 + 2] = fmaf(fa.z, ln_k2, ba[q4 * 4 + 2]); bg[q4 * 4 + 2] = fmaf(fg.z, ln_k2, bg[q4 * 4 + 2]);
              ba[q4 * 4 + 3] = fmaf(fa.w, ln_k2, ba[q4 * 4 + 3]); bg[q4 * 4 + 3] = fmaf(fg.w, ln_k2, bg[q4 * 4 + 3]);
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {  // ln_rstd == 1 without the fold
            const float a = fmaf(__uint_as_float(ra[j]), ln_rstd, ba[j]);
            const float gt = fmaf(__uint_as_float(rg[j]), ln_rstd, bg[j]);
            v[j] = a * gelu_erf(gt);
          }
        }

        if (p.has_res) {
          mbar_wait(&my_res_bar[buf], buf_parity);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const uint4 rv = *reinterpret_cast<const uint4*>(sbuf + lane * 64 + ((ch ^ ((lane >> 1) & 3)) * 16));
            const float2 f0 = unpack_bf16x2(rv.x), f1 = unpack_bf16x2(rv.y), f2 = unpack_bf16x2(rv.z),
                         f3 = unpack_bf16x2(rv.w);
            v[ch * 8 + 0] += f0.x; v[ch * 8 + 1] += f0.y; v[ch * 8 + 2] += f1.x; v[ch * 8 + 3] += f1.y;
            v[ch * 8 + 4] += f2.x; v[ch * 8 + 5] += f2.y; v[ch * 8 + 6] += f3.x; v[ch * 8 + 7] += f3.y;
          }
        }
        if (p.film_scale != nullptr && col0 + 32 <= n_out) {
          const float4* fs = reinterpret_cast<const float4*>(p.film_scale + static_cast<size_t>(img) * p.film_ld + col0);
          const float4* fb = reinterpret_cast<const float4*>(p.film_shift + static_cast<size_t>(img) * p.film_ld + col0);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 a4 = __ldg(fs + q4), b4 = __ldg(fb + q4);
            v[q4 * 4 + 0] = fmaf(v[q4 * 4 + 0], a4.x, b4.x);
            v[q4 * 4 + 1] = fmaf(v[q4 * 4 + 1], a4.y, b4.y);
            v[q4 * 4 + 2] = fmaf(v[q4 * 4 + 2], a4.z, b4.z);
            v[q4 * 4 + 3] = fmaf(v[q4 * 4 + 3], a4.w, b4.w);
          }
        }
        if (p.stats_out != nullptr && col0 < n_out) {
          // statistics of the fp32 values (the stored ones are their bf16 roundings: zero-mean noise of 2^-9 relative
          // size per element, which changes mean and variance of a >= 320-wide row by less than 1e-4 relative)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            st_sum += v[j];
            st_sq = fmaf(v[j], v[j], st_sq);
          }
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint4 o;
          o.x = pack_bf16x2(v[ch * 8 + 0], v[ch * 8 + 1]);
          o.y = pack_bf16x2(v[ch * 8 + 2], v[ch * 8 + 3]);
          o.z = pack_bf16x2(v[ch * 8 + 4], v[ch * 8 + 5]);
          o.w = pack_bf16x2(v[ch * 8 + 6], v[ch * 8 + 7]);
          *reinterpret_cast<uint4*>(sbuf + lane * 64 + ((ch ^ ((lane >> 1) & 3)) * 16)) = o;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (col0 < n_out) tma_store_4d(&mapOut, sbuf, col0, sx, sy, sn);
          tma_store_commit();
        }
      }
      if (p.stats_out != nullptr && ln == 0 && x0 + lx < p.rows_total)
        p.stats_out[static_cast<size_t>(x0 + lx) * p.tiles_n + n_tile] = make_float2(st_sum, st_sq);
      // accumulator fully read -> hand the TMEM stage back to the MMA warp (published segments: already done above)
      tc_fence_before();
      __syncwarp();
      if (part_rd == nullptr && lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (lane == 0) tma_store_wait_all0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct ConvGeom {
  int Nimg, H, W;     // OUTPUT geometry (pixels)
  int TW, TH, TN;
};

static void pick_tile(int Nimg, int H, int W, int* TW, int* TH, int* TN) {
  // 128 output pixels per tile as TN images x TH rows x TW pixels; prefer exact cover along x.
  int tw = 128;
  while (tw > 1 && tw > W) tw >>= 1;            // largest power of two <= W (<=128)
  if (H > 1 && W % tw != 0) {                   // e.g. W = 96 -> 32, 48 -> 16, 24 -> 8, 12 -> 4
    int cand = tw;
    while (cand > 1 && W % cand != 0) cand >>= 1;
    if (cand >= 4) tw = cand;
  }
  int rest = 128 / tw;
  int th = 1;
  while (th * 2 <= rest && th * 2 <= H) th <<= 1;
  if (H % th != 0) {
    int cand = th;
    while (cand > 1 && H % cand != 0) cand >>= 1;
    th = cand;
  }
  *TW = tw;
  *TH = th;
  *TN = 128 / (tw * th);
  (void)Nimg;
}

// Weight-stationary tiles for the K <= 320 linears whose N is a multiple of 128 (measured in round 2: q,k,v,q_ref at
// M = 32768 38.5 -> 36.1 us, step 13.26 -> 13.08 ms). MVD_GEMM_WS=0 switches them off for A/B runs.
static bool ws_enabled() {
  static const bool on = [] {
    const char* e = getenv("MVD_GEMM_WS");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

template <int BN, bool S2>
static int launch_one(const CUtensorMap& mA, const CUtensorMap& mA2, const CUtensorMap& mB, const CUtensorMap& mO,
                      const CUtensorMap& mR, const GemmArgs& args, cudaStream_t stream) {
  using L = SmemLayout<BN>;
  // per call: cheap, and correct for every device a process may touch
  MVD_CUDA(cudaFuncSetAttribute(gemm_conv_kernel<BN, S2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  int units = sm_count();  // one CTA per SM
  if (args.tiles_total < units) units = args.tiles_total;
  MVD_CUDA(launch_pdl(gemm_conv_kernel<BN, S2>, dim3(units), dim3(GEMM_THREADS), L::TOTAL, stream, mA, mA2, mB, mO, mR,
                      args));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

static int launch_ws(const CUtensorMap& mA, const CUtensorMap& mA2, const CUtensorMap& mB, const CUtensorMap& mO,
                     const CUtensorMap& mR, const GemmArgs& args, cudaStream_t stream) {
  using L = SmemLayout<128, true>;
  MVD_CUDA(cudaFuncSetAttribute(gemm_conv_kernel<128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                L::TOTAL));
  // every CTA keeps one column tile: grid = tiles_n * (row groups that fit on the machine)
  const int tiles_m = args.tiles_total / args.tiles_n;
  int groups = sm_count() / args.tiles_n;
  if (groups > tiles_m) groups = tiles_m;
  MVD_CUDA(launch_pdl(gemm_conv_kernel<128, false, true>, dim3(groups * args.tiles_n), dim3(GEMM_THREADS), L::TOTAL,
                      stream, mA, mA2, mB, mO, mR, args));
  MVD_CUDA(cudaGetLastError());
  count_launches(1);
  return MVD_OK;
}

static double bn_cost(int N, int M_tiles, int bn, int units) {
  const int tn = (N + bn - 1) / bn;
  const long tiles = static_cast<long>(tn) * M_tiles;
  const long waves = (tiles + units - 1) / units;
  // cost ~ waves * per-tile time (prop. to bn) ; small fixed per-tile overhead favours larger tiles
  return static_cast<double>(waves) * (bn + 24.0);
}

static int pick_bn(int N, int M_tiles, bool geglu) {
  // Candidate tile widths; choose the one with the least padded work, ties -> fewer waves imbalance.
  const int cands_plain[] = {256, 160, 128, 64};
  const int cands_geglu[] = {256, 128, 64};
  const int* cands = geglu ? cands_geglu : cands_plain;
  const int nc = geglu ? 3 : 4;
  int best = 64;
  double best_cost = 1e30;
  for (int i = 0; i < nc; ++i) {
    const double cost = bn_cost(N, M_tiles, cands[i], sm_count());
    if (cost < best_cost) {
      best_cost = cost;
      best = cands[i];
    }
  }
  return best;
}

// The scheduling decision for one problem: pixel tile shape, tile width, weight-stationary or streaming, stream-K.
struct GemmPlan {
  int TW, TH, TN, tiles_m;
  int bn;
  bool ws;       // weight-stationary 128-wide tiles (grid must then be a multiple of the column-tile count)
  int sk_first;  // stream-K (needs a workspace): tiles [sk_first, tiles) are shared k-block-wise by sk_ctas CTAs
  int sk_ctas;   // 0: no stream-K
  int sk_slots;
};

constexpr int64_t SPLITK_TICKET_BYTES = 4096 * sizeof(unsigned int);
constexpr int SPLITK_BN = 64;

static int streamk_mode() {  // MVD_STREAMK: 0 = off, 1 = launches smaller than the machine only, 2 (default) = + tails
  static const int v = [] {
    const char* e = getenv("MVD_STREAMK");
    return e == nullptr ? 2 : atoi(e);
  }();
  return v;
}

// Stream-K for the part of a launch that does not fill the machine: the last, partial wave of `rem` tiles (or all tiles
// of a launch with fewer tiles than SMs). Whole tiles would keep `rem` SMs busy for k_blocks k-block steps; sharing
// the rem * k_blocks units evenly keeps every SM busy for `per` steps, at the price of the partial exchange (an fp32
// tile out per segment, `slots` of them back in through L2 for the segment that arrives last, a ticket). Costs are in
// k-block steps of a 64-wide tile: a step moves 16 KB of A and bn / 8 KB of B into the SM (the K-long launches this is
// for are L2->SM bound: profiles/r1_tma_bw.txt), a tile has ~16 steps of fixed cost, the exchange 8 + 3 per slot and
// 64 columns (fitted to profiles/r2_streamk_time.txt, where the model picks the fastest width at every shape).
struct StreamK {
  int first, ctas, slots;
  double cost;  // of the whole launch
};
static StreamK eval_streamk(int tiles, int k_blocks, int bn, int64_t ws_bytes) {
  const int sms = sm_count();
  const double w = (16.0 + bn / 8.0) / 24.0;
  const int first = tiles / sms * sms;
  const int rem = tiles - first;
  StreamK r{tiles, 0, 0, (first / sms) * k_blocks * w + 16.0 + (rem > 0 ? k_blocks * w : 0.0)};
  const int mode = streamk_mode();
  if (rem == 0 || mode <= 0 || ws_bytes <= SPLITK_TICKET_BYTES || (first > 0 && mode < 2)) return r;
  const int units = rem * k_blocks;
  constexpr int kMinUnits = 12;  // k-block steps per CTA below which the fixed costs dominate
  const int ctas = units / kMinUnits < sms ? units / kMinUnits : sms;
  if (ctas <= rem || static_cast<int64_t>(units) * ctas >= (1LL << 31)) return r;
  const int per = (units + ctas - 1) / ctas;
  // segments per tile, exactly as WorkIter assigns them (CTA i owns units [floor(i T / G), floor((i + 1) T / G)))
  int slots = 1;
  for (int t = 0; t < rem; ++t) {
    const int c_first = ((t * k_blocks + 1) * ctas - 1) / units;
    const int c_last = ((t * k_blocks + k_blocks) * ctas - 1) / units;
    if (c_last - c_first + 1 > slots) slots = c_last - c_first + 1;
  }
  if (rem * 4 > static_cast<int>(SPLITK_TICKET_BYTES / sizeof(unsigned int)) ||
      SPLITK_TICKET_BYTES + static_cast<int64_t>(rem) * slots * BM * bn * 4 > ws_bytes)
    return r;
  const double whole = k_blocks * w + 16.0;
  const double shared = per * w + 16.0 + 8.0 + 3.0 * slots * (bn / 64.0);
  if (shared < 0.85 * whole) {
    r.first = first;
    r.ctas = ctas;
    r.slots = slots;
    r.cost = (first / sms) * k_blocks * w + shared;
  }
  return r;
}
static GemmPlan plan_gemm(int Nimg, int H, int W, int Cin, int Cout, int ntaps, int stride, int geglu, int force_bn,
                          bool two_source, int64_t ws_bytes = 0) {
  GemmPlan pl;
  pick_tile(Nimg, H, W, &pl.TW, &pl.TH, &pl.TN);
  pl.tiles_m = ((W + pl.TW - 1) / pl.TW) * ((H + pl.TH - 1) / pl.TH) * ((Nimg + pl.TN - 1) / pl.TN);
  pl.bn = force_bn > 0 ? force_bn : pick_bn(Cout, pl.tiles_m, geglu != 0);
  // weight-stationary: 1-tap, K <= 320, N a multiple of 128 (GEGLU: only with the 128-wide interleave), enough row
  // tiles that every CTA amortises its resident weight tile over several of them
  pl.ws = ws_enabled() && ntaps == 1 && stride == 1 && Cin <= WS_MAX_KBLOCKS * BK && Cout % 128 == 0 &&
          (force_bn == 0 || force_bn == 128) && (!geglu || force_bn == 128) && !two_source &&
          Cout / 128 <= sm_count() && pl.tiles_m * (Cout / 128) >= 4 * sm_count();
  if (pl.ws) pl.bn = 128;
  auto tiles_of = [&](int bn) { return ((Cout + bn - 1) / bn) * pl.tiles_m; };
  pl.sk_first = tiles_of(pl.bn);
  pl.sk_ctas = pl.sk_slots = 0;
  if (!pl.ws && !geglu && stride == 1) {
    const int k_blocks = ntaps * (Cin / 64);
    StreamK best = eval_streamk(tiles_of(pl.bn), k_blocks, pl.bn, ws_bytes);
    // launches of less than two waves of 64-wide tiles: with the k-blocks shared, wave quantisation no longer favours
    // narrow tiles, so the width is chosen again by the cost of the whole launch
    if (force_bn == 0 && ws_bytes > SPLITK_TICKET_BYTES && streamk_mode() > 0 && tiles_of(64) < 2 * sm_count()) {
      const int cands[3] = {64, 128, 160};
      best.cost = 1e30;
      for (int i = 0; i < 3; ++i) {
        const StreamK c = eval_streamk(tiles_of(cands[i]), k_blocks, cands[i], ws_bytes);
        if (c.cost < best.cost) {
          best = c;
          pl.bn = cands[i];
        }
      }
    }
    pl.sk_first = best.first;
    pl.sk_ctas = best.ctas;
    pl.sk_slots = best.slots;
  }
  return pl;
}

// Generic launcher. x: NHWC-like activation described as (C, Wd, Hd, Nd) with element strides.
struct OperandA {
  const void* ptr;
  int C;                 // channels of this source
  int64_t pix_stride;    // elements between consecutive pixels (>= C)
};

static int run_gemm_conv(const OperandA& a1, const OperandA* a2, const void* w, int64_t ldw, const void* bias,
                         const float* img_bias, int img_bias_ld, int rows_per_img, const void* residual,
                         int64_t res_pix_stride, void* out, int64_t out_pix_stride, int Nimg, int H, int W /*output*/,
                         int Cout, int ntaps, int stride, int geglu, int force_bn, const mvd_gemm_extras* ex,
                         cudaStream_t stream) {
  const int Cin = a1.C + (a2 ? a2->C : 0);
  MVD_CHECK(Cin % 64 == 0 && a1.C % 64 == 0, "gemm/conv: K (=%d, first source %d) must be a multiple of 64", Cin,
            a1.C);
  MVD_CHECK(Cout % 32 == 0, "gemm/conv: N (=%d) must be a multiple of 32", Cout);
  MVD_CHECK(!(stride == 2 && (a2 || ntaps != 9)), "stride-2 supports single-source 3x3 only");
  MVD_CHECK(!geglu || (force_bn > 0 && Cout % force_bn == 0 && !residual && !img_bias),
            "geglu needs an explicit tile width dividing N, and no residual");
  MVD_CHECK((reinterpret_cast<uintptr_t>(a1.ptr) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(out) & 15) == 0,
            "gemm/conv: pointers must be 16-byte aligned");
  MVD_CHECK(a1.pix_stride % 8 == 0 && out_pix_stride % 8 == 0 && ldw % 8 == 0,
            "gemm/conv: strides must be multiples of 8 elements");
  MVD_CHECK((reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm/conv: bias must be 16-byte aligned");
  MVD_CHECK(img_bias == nullptr || ((reinterpret_cast<uintptr_t>(img_bias) & 15) == 0 && img_bias_ld % 4 == 0),
            "gemm/conv: per-image bias must be 16-byte aligned with a row stride multiple of 4");

  GemmArgs g;
  memset(&g, 0, sizeof(g));
  const bool plain_ex = ex == nullptr || (ex->ln_stats == nullptr && ex->stats_out == nullptr);
  const GemmPlan pl = plan_gemm(Nimg, H, W, Cin, Cout, ntaps, stride, geglu, force_bn, a2 != nullptr,
                                (ex != nullptr && ex->workspace != nullptr && plain_ex) ? ex->workspace_bytes : 0);
  const int TW = pl.TW, TH = pl.TH, TN = pl.TN;
  g.TW = TW; g.TH = TH; g.TN = TN;
  g.tiles_x = (W + TW - 1) / TW;
  g.tiles_y = (H + TH - 1) / TH;
  const int tiles_img = (Nimg + TN - 1) / TN;
  const int tiles_m = g.tiles_x * g.tiles_y * tiles_img;
  g.bx = TW < 32 ? TW : 32;
  g.by = (32 / g.bx) < TH ? (32 / g.bx) : TH;
  const int bn_img = 32 / (g.bx * g.by);
  g.ntaps = ntaps;
  g.kc_per_tap = Cin / 64;
  g.kc_a1 = a1.C / 64;
  g.cin = Cin;
  g.c_s2 = a1.C;
  g.N = Cout;
  g.geglu = geglu;
  g.has_res = residual != nullptr;
  g.rows_per_img = rows_per_img > 0 ? rows_per_img : (1 << 30);
  g.img_bias = img_bias;
  g.img_bias_ld = img_bias_ld;
  g.img_max = rows_per_img > 0 ? (W - 1) / rows_per_img : Nimg - 1;
  g.bias = static_cast<const __nv_bfloat16*>(bias);

  const int BN = pl.bn;
  g.tiles_n = (Cout + BN - 1) / BN;
  g.tiles_total = g.tiles_n * tiles_m;
  g.rows_total = W;
  g.sk_first = g.tiles_total;
  if (pl.sk_ctas > 0) {
    MVD_CHECK((reinterpret_cast<uintptr_t>(ex->workspace) & 15) == 0, "gemm: workspace must be 16-byte aligned");
    g.ws_tickets = static_cast<unsigned int*>(ex->workspace);
    g.ws_partial = reinterpret_cast<float4*>(static_cast<char*>(ex->workspace) + SPLITK_TICKET_BYTES);
    g.sk_first = pl.sk_first;
    g.sk_ctas = pl.sk_ctas;
    g.sk_slots = pl.sk_slots;
    g.sk_units = (g.tiles_total - pl.sk_first) * ntaps * (Cin / 64);
  }
  if (ex != nullptr) {
    const bool linear_mode = ntaps == 1 && H == 1 && Nimg == 1;
    if (ex->ln_stats != nullptr) {
      MVD_CHECK(linear_mode && a2 == nullptr && ex->ln_parts > 0 && ex->ln_colsum != nullptr,
                "gemm: the LayerNorm fold needs a single-source linear, ln_parts > 0 and column sums");
      MVD_CHECK(geglu || (img_bias != nullptr && bias == nullptr),
                "gemm: with the LayerNorm fold the constant term (W.beta + bias) is passed as the fp32 row-group bias");
      MVD_CHECK(((reinterpret_cast<uintptr_t>(ex->ln_stats) & 7) | (reinterpret_cast<uintptr_t>(ex->ln_colsum) & 15)) == 0,
                "gemm: ln_stats must be 8-byte and ln_colsum 16-byte aligned");
      g.ln_stats = reinterpret_cast<const float2*>(ex->ln_stats);
      g.ln_colsum = ex->ln_colsum;
      g.ln_parts = ex->ln_parts;
      g.ln_inv_k = 1.0f / static_cast<float>(Cin);
      g.ln_eps = ex->ln_eps;
    }
    if (ex->stats_out != nullptr) {
      MVD_CHECK(linear_mode && !geglu, "gemm: row statistics are produced by plain linears only");
      MVD_CHECK(ex->stats_parts == g.tiles_n, "gemm: stats_out was sized for %d column tiles, this launch has %d "
                "(ask mvd_gemm_plan)", ex->stats_parts, g.tiles_n);
      MVD_CHECK((reinterpret_cast<uintptr_t>(ex->stats_out) & 7) == 0, "gemm: stats_out must be 8-byte aligned");
      g.stats_out = reinterpret_cast<float2*>(ex->stats_out);
    }
    if (ex->film_scale != nullptr) {
      MVD_CHECK(!geglu && ex->film_shift != nullptr && ex->film_ld % 4 == 0 &&
                    ((reinterpret_cast<uintptr_t>(ex->film_scale) | reinterpret_cast<uintptr_t>(ex->film_shift)) & 15) == 0,
                "gemm: FiLM needs 16-byte aligned fp32 scale and shift rows with a stride multiple of 4");
      g.film_scale = ex->film_scale;
      g.film_shift = ex->film_shift;
      g.film_ld = ex->film_ld;
    }
  }

  CUtensorMap mA, mA2, mB, mO, mR;
  // --- A maps
  if (stride == 2) {
    // input is [Nimg, 2H, 2W, C]; view (2C, W, 2, H, Nimg)
    const uint64_t ps = static_cast<uint64_t>(a1.pix_stride);
    const uint64_t dims[5] = {static_cast<uint64_t>(2 * a1.C), static_cast<uint64_t>(W), 2, static_cast<uint64_t>(H),
                              static_cast<uint64_t>(Nimg)};
    MVD_CHECK(a1.pix_stride == a1.C, "stride-2 conv needs a dense NHWC input");
    const uint64_t strides[4] = {2 * ps * 2, 2 * static_cast<uint64_t>(W) * ps * 2,
                                 2 * 2 * static_cast<uint64_t>(W) * ps * 2,
                                 static_cast<uint64_t>(2 * H) * 2 * W * ps * 2};
    const uint32_t box[5] = {64, static_cast<uint32_t>(TW), 1, static_cast<uint32_t>(TH), static_cast<uint32_t>(TN)};
    if (int e = make_tmap_bf16(&mA, a1.ptr, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    mA2 = mA;
  } else {
    auto mk = [&](CUtensorMap* m, const OperandA& a) -> int {
      const uint64_t ps = static_cast<uint64_t>(a.pix_stride);
      const uint64_t dims[4] = {static_cast<uint64_t>(a.C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                                static_cast<uint64_t>(Nimg)};
      const uint64_t strides[3] = {ps * 2, static_cast<uint64_t>(W) * ps * 2,
                                   static_cast<uint64_t>(H) * W * ps * 2};
      const uint32_t box[4] = {64, static_cast<uint32_t>(TW), static_cast<uint32_t>(TH), static_cast<uint32_t>(TN)};
      return make_tmap_bf16(m, a.ptr, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    };
    if (int e = mk(&mA, a1)) return e;
    if (a2) {
      if (int e = mk(&mA2, *a2)) return e;
    } else {
      mA2 = mA;
    }
  }
  // --- B map: [Cout, ntaps*Cin] row-major
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(ntaps) * Cin, static_cast<uint64_t>(Cout)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(BN)};
    if (int e = make_tmap_bf16(&mB, w, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  // --- output / residual maps (per-warp slab boxes of 32 pixels x 32 channels, 64-B swizzle)
  {
    const int n_out = geglu ? Cout / 2 : Cout;
    auto mk = [&](CUtensorMap* m, const void* ptr, int64_t pstride) -> int {
      const uint64_t ps = static_cast<uint64_t>(pstride);
      const uint64_t dims[4] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                                static_cast<uint64_t>(Nimg)};
      const uint64_t strides[3] = {ps * 2, static_cast<uint64_t>(W) * ps * 2,
                                   static_cast<uint64_t>(H) * W * ps * 2};
      const uint32_t box[4] = {32, static_cast<uint32_t>(g.bx), static_cast<uint32_t>(g.by),
                               static_cast<uint32_t>(bn_img)};
      return make_tmap_bf16(m, ptr, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
    };
    if (int e = mk(&mO, out, out_pix_stride)) return e;
    if (residual) {
      MVD_CHECK(res_pix_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
                "gemm/conv: residual must be 16-byte aligned with stride %% 8 == 0");
      if (int e = mk(&mR, residual, res_pix_stride)) return e;
    } else {
      mR = mO;
    }
  }

  if (pl.ws) return launch_ws(mA, mA2, mB, mO, mR, g, stream);
  if (pl.sk_ctas > 0) {
    // whole-tile work (if any) is spread over every SM; a launch that is stream-K only needs just its CTAs
    const int grid = pl.sk_first > 0 ? sm_count() : pl.sk_ctas;
#define MVD_LAUNCH_SK(bn)                                                                                             \
  case bn: {                                                                                                          \
    using L = SmemLayout<bn>;                                                                                         \
    MVD_CUDA(cudaFuncSetAttribute(gemm_conv_kernel<bn, false, false, true>,                                           \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));                            \
    MVD_CUDA(launch_pdl(gemm_conv_kernel<bn, false, false, true>, dim3(grid), dim3(GEMM_THREADS), L::TOTAL, stream,   \
                        mA, mA2, mB, mO, mR, g));                                                                     \
    break;                                                                                                            \
  }
    switch (BN) {
      MVD_LAUNCH_SK(64)
      MVD_LAUNCH_SK(128)
      MVD_LAUNCH_SK(160)
      MVD_LAUNCH_SK(256)
      default:
        set_error("gemm/conv: unsupported tile width %d", BN);
        return MVD_ERR_INVALID;
    }
#undef MVD_LAUNCH_SK
    MVD_CUDA(cudaGetLastError());
    count_launches(1);
    return MVD_OK;
  }
#define MVD_LAUNCH_BN(bn)                                                               \
  case bn:                                                                              \
    return stride == 2 ? launch_one<bn, true>(mA, mA2, mB, mO, mR, g, stream)           \
                       : launch_one<bn, false>(mA, mA2, mB, mO, mR, g, stream);
  switch (BN) {
    MVD_LAUNCH_BN(64)
    MVD_LAUNCH_BN(128)
    MVD_LAUNCH_BN(160)
    MVD_LAUNCH_BN(256)
    default:
      set_error("gemm/conv: unsupported tile width %d", BN);
      return MVD_ERR_INVALID;
  }
#undef MVD_LAUNCH_BN
}

}  // namespace mvd

extern "C" {

int mvd_gemm_plan(int n_img, int h_out, int w_out, int c_in, int c_out, int ntaps, int stride, int geglu, int tile_n,
                  int* bn, int* weight_stationary, int* grid) {
  using namespace mvd;
  MVD_CHECK(n_img > 0 && h_out > 0 && w_out > 0 && c_in > 0 && c_in % 64 == 0 && c_out > 0 && c_out % 32 == 0 &&
                (ntaps == 1 || ntaps == 9) && (stride == 1 || stride == 2),
            "gemm_plan: unsupported problem");
  const GemmPlan pl = plan_gemm(n_img, h_out, w_out, c_in, c_out, ntaps, stride, geglu, tile_n, false);
  const int units = ((c_out + pl.bn - 1) / pl.bn) * pl.tiles_m;
  if (bn) *bn = pl.bn;
  if (weight_stationary) *weight_stationary = pl.ws ? 1 : 0;
  if (grid) {
    *grid = units < sm_count() ? units : sm_count();
    if (pl.ws) {
      const int tn = c_out / 128;
      const int groups = sm_count() / tn < pl.tiles_m ? sm_count() / tn : pl.tiles_m;
      *grid = groups * tn;
    }
  }
  return MVD_OK;
}

int mvd_gemm_plan_streamk(int n_img, int h_out, int w_out, int c_in, int c_out, int ntaps, int stride, int tile_n,
                          int64_t workspace_bytes, int* bn, int* tiles, int* k_blocks, int* sk_first, int* sk_ctas,
                          int* sk_slots) {
  using namespace mvd;
  MVD_CHECK(n_img > 0 && h_out > 0 && w_out > 0 && c_in > 0 && c_in % 64 == 0 && c_out > 0 && c_out % 32 == 0 &&
                (ntaps == 1 || ntaps == 9) && (stride == 1 || stride == 2),
            "gemm_plan: unsupported problem");
  const GemmPlan pl = plan_gemm(n_img, h_out, w_out, c_in, c_out, ntaps, stride, /*geglu=*/0, tile_n, false,
                                workspace_bytes);
  if (bn) *bn = pl.bn;
  if (tiles) *tiles = ((c_out + pl.bn - 1) / pl.bn) * pl.tiles_m;
  if (k_blocks) *k_blocks = ntaps * (c_in / 64);
  if (sk_first) *sk_first = pl.sk_first;
  if (sk_ctas) *sk_ctas = pl.sk_ctas;
  if (sk_slots) *sk_slots = pl.sk_slots;
  return MVD_OK;
}

int mvd_linear_ex_bf16(const void* a, int64_t lda, int k1, const void* a2, int64_t lda2, int k2, const void* w,
                       int64_t ldw, const void* bias, const float* row_group_bias, int row_group_bias_ld,
                       int rows_per_group, const void* residual, int64_t ldr, void* out, int64_t ldo, int M, int N,
                       int geglu, int tile_n, const mvd_gemm_extras* extras, void* stream) {
  using namespace mvd;
  MVD_CHECK(M > 0 && N > 0 && k1 > 0, "linear: empty problem M=%d N=%d K=%d", M, N, k1);
  OperandA s1{a, k1, lda};
  OperandA s2{a2, k2, lda2};
  return run_gemm_conv(s1, (a2 && k2 > 0) ? &s2 : nullptr, w, ldw, bias, row_group_bias, row_group_bias_ld,
                       rows_per_group, residual, ldr, out, ldo, /*Nimg=*/1, /*H=*/1, /*W=*/M, N, /*ntaps=*/1,
                       /*stride=*/1, geglu, tile_n, extras, static_cast<cudaStream_t>(stream));
}

int mvd_linear_bf16(const void* a, int64_t lda, int k1, const void* a2, int64_t lda2, int k2, const void* w,
                    int64_t ldw, const void* bias, const float* row_group_bias, int row_group_bias_ld,
                    int rows_per_group, const void* residual, int64_t ldr, void* out, int64_t ldo, int M, int N,
                    int geglu, int tile_n, void* stream) {
  return mvd_linear_ex_bf16(a, lda, k1, a2, lda2, k2, w, ldw, bias, row_group_bias, row_group_bias_ld, rows_per_group,
                            residual, ldr, out, ldo, M, N, geglu, tile_n, nullptr, stream);
}

int mvd_conv3x3_ex_bf16(const void* x, int cin1, const void* x2, int cin2, const void* w, const void* bias,
                        const float* img_bias, int img_bias_ld, const void* residual, void* out, int n_img, int h_out,
                        int w_out, int c_out, int stride, int tile_n, const mvd_gemm_extras* extras, void* stream) {
  using namespace mvd;
  MVD_CHECK(n_img > 0 && h_out > 0 && w_out > 0, "conv3x3: empty problem");
  MVD_CHECK(stride == 1 || stride == 2, "conv3x3: stride must be 1 or 2");
  OperandA s1{x, cin1, cin1};
  OperandA s2{x2, cin2, cin2};
  const int cin = cin1 + ((x2 && cin2 > 0) ? cin2 : 0);
  return run_gemm_conv(s1, (x2 && cin2 > 0) ? &s2 : nullptr, w, static_cast<int64_t>(9) * cin, bias, img_bias,
                       img_bias_ld, /*rows_per_group=*/0, residual, c_out, out, c_out, n_img, h_out, w_out, c_out,
                       /*ntaps=*/9, stride, /*geglu=*/0, tile_n, extras, static_cast<cudaStream_t>(stream));
}

int mvd_conv3x3_bf16(const void* x, int cin1, const void* x2, int cin2, const void* w, const void* bias,
                     const float* img_bias, int img_bias_ld, const void* residual, void* out, int n_img, int h_out,
                     int w_out, int c_out, int stride, int tile_n, void* stream) {
  return mvd_conv3x3_ex_bf16(x, cin1, x2, cin2, w, bias, img_bias, img_bias_ld, residual, out, n_img, h_out, w_out,
                             c_out, stride, tile_n, nullptr, stream);
}

}  // extern "C"
