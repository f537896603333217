// Micro-benchmark: per-SM and chip-wide throughput of TMA tensor loads and stores for the box shapes the GEMM kernel
// uses, to decide whether the small-K linears are bound by L2->SM operand traffic, by the 64-byte-row output
// stores of the epilogue, or by neither.
//
//   loads : 128 rows x 128 B (one A k-block, SWIZZLE_128B) through a 4-deep smem ring, one issuing thread per CTA
//   stores: 32 rows x 64 B (the current epilogue chunk, SWIZZLE_64B) and 32 rows x 128 B (SWIZZLE_128B), four
//           warps per CTA, up to 4 bulk groups in flight per warp, into a matrix with a 2560-byte row pitch
//
// Each is run with 1, 37, 74 and 148 CTAs (one per SM). Output: GB/s, bytes/clk/SM at the measured SM clock.
// --mix: loads and stores together; --depth: load ring depth (4/8/12 boxes in flight) x private/shared source windows.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../mvd_b200/csrc tma_bw.cu \
//             ../../mvd_b200/csrc/host_common.cu -o tma_bw -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc.cuh"
#include "host_common.h"
#include "../../include/mvd_b200.h"

using namespace mvd;

constexpr int ROWS = 32768, COLS = 1280;  // bf16 matrix, 2560-byte pitch (the q,k,v,q_ref projection output)

__global__ void __launch_bounds__(128, 1)
load_bw_kernel(const __grid_constant__ CUtensorMap map, int iters, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[4];
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    // 4 loads in flight: wait for box i-4 before re-using its slot
    for (int i = 0; i < iters + 4; ++i) {
      const int s = i & 3;
      if (i >= 4) mbar_wait(&full[s], ((i - 4) >> 2) & 1);
      if (i < iters) {
        mbar_arrive_expect_tx(&full[s], 128 * 128);
        const int box = blockIdx.x * iters + i;  // distinct boxes: (ROWS/128) x (COLS/64) = 256 x 20 of them
        tma_load_2d(smem + s * 16384, &map, &full[s], (box % (COLS / 64)) * 64, ((box / (COLS / 64)) % (ROWS / 128)) * 128);
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

template <int ROW_BYTES>
__global__ void __launch_bounds__(128, 1)
store_bw_kernel(const __grid_constant__ CUtensorMap map, int iters, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  constexpr int BOX_BYTES = 32 * ROW_BYTES;
  constexpr int BOX_COLS = ROW_BYTES / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* my = smem + warp * 4 * BOX_BYTES;
  for (int i = lane; i < 4 * BOX_BYTES / 4; i += 32) reinterpret_cast<uint32_t*>(my)[i] = 0x3f803f80u;
  fence_proxy_async_smem();
  __syncthreads();
  const long long t0 = clock64();
  if (lane == 0) {
    // same traversal as the epilogue: a warp owns 32 rows of a 128-row tile and walks the columns of its tile row
    const int chunks_per_row = COLS / BOX_COLS;
    for (int i = 0; i < iters; ++i) {
      asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      const int idx = blockIdx.x * iters + i;
      const int m_tile = (idx / chunks_per_row) % (ROWS / 128);
      tma_store_2d(&map, my + (i & 3) * BOX_BYTES, (idx % chunks_per_row) * BOX_COLS, m_tile * 128 + warp * 32);
      tma_store_commit();
    }
    tma_store_wait_all0();
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}


// Loads and stores together: warp 0 runs the TMA load ring, warps 1-4 store 32-row x 64-B chunks either with TMA
// (mode 1) or with coalesced st.global.v4 from registers, 8 rows x 64 B per warp instruction (mode 2); mode 0: loads
// only, mode 3: st.global only. Do the two TMA directions share one per-SM budget? Does the LSU path add to it?
__global__ void __launch_bounds__(160, 1)
mix_kernel(const __grid_constant__ CUtensorMap map_ld, const __grid_constant__ CUtensorMap map_st, __nv_bfloat16* out,
           int iters, int mode, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  uint8_t* stage = smem + 65536 + (warp > 0 ? warp - 1 : 0) * 8192;
  if (warp > 0)
    for (int i = lane; i < 8192 / 4; i += 32) reinterpret_cast<uint32_t*>(stage)[i] = 0x3f803f80u;
  fence_proxy_async_smem();
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0 && mode != 3) {
      for (int i = 0; i < iters + 4; ++i) {
        const int s = i & 3;
        if (i >= 4) mbar_wait(&full[s], ((i - 4) >> 2) & 1);
        if (i < iters) {
          mbar_arrive_expect_tx(&full[s], 128 * 128);
          const int box = blockIdx.x * iters + i;
          tma_load_2d(smem + s * 16384, &map_ld, &full[s], (box % (COLS / 64)) * 64,
                      ((box / (COLS / 64)) % (ROWS / 128)) * 128);
        }
      }
      cycles[blockIdx.x] = clock64() - t0;
    }
  } else if (mode == 1) {
    if (lane == 0) {
      for (int i = 0; i < iters; ++i) {
        asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
        const int idx = blockIdx.x * iters + i;
        tma_store_2d(&map_st, stage + (i & 3) * 2048, (idx % (COLS / 32)) * 32,
                     ((idx / (COLS / 32)) % (ROWS / 128)) * 128 + (warp - 1) * 32);
        tma_store_commit();
      }
      tma_store_wait_all0();
      if (warp == 1) cycles[256 + blockIdx.x] = clock64() - t0;
    }
  } else if (mode >= 2) {
    const uint4 v = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    for (int i = 0; i < iters; ++i) {
      const int idx = blockIdx.x * iters + i;
      const int row0 = ((idx / (COLS / 32)) % (ROWS / 128)) * 128 + (warp - 1) * 32;
      const int col0 = (idx % (COLS / 32)) * 32;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int row = row0 + r * 8 + (lane >> 2);
        *reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * COLS + col0 + (lane & 3) * 8) = v;
      }
    }
    if (warp == 1 && lane == 0) cycles[256 + blockIdx.x] = clock64() - t0;
  }
}


// Load throughput vs ring depth and sharing: DEPTH boxes of 16 KB in flight per CTA; `shared_rows` > 0 makes every
// CTA walk the same `shared_rows`-row window (weight-like: the same lines are wanted by all SMs at about the same
// time), 0 gives every CTA its own boxes (activation-like). Separates a per-SM ingest limit from an L2-side one.
template <int DEPTH>
__global__ void __launch_bounds__(128, 1)
load_depth_kernel(const __grid_constant__ CUtensorMap map, int iters, int shared_rows, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[DEPTH];
  if (threadIdx.x == 0) {
    for (int s = 0; s < DEPTH; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < iters + DEPTH; ++i) {
      const int s = i % DEPTH;
      if (i >= DEPTH) mbar_wait(&full[s], ((i - DEPTH) / DEPTH) & 1);
      if (i < iters) {
        mbar_arrive_expect_tx(&full[s], 128 * 128);
        const int box = shared_rows > 0 ? i : blockIdx.x * iters + i;
        const int row_tiles = shared_rows > 0 ? shared_rows / 128 : ROWS / 128;
        tma_load_2d(smem + s * 16384, &map, &full[s], (box % (COLS / 64)) * 64, ((box / (COLS / 64)) % row_tiles) * 128);
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("%s, %d SMs, nominal %d MHz\n", prop.name, prop.multiProcessorCount, clk_khz / 1000);
  void* buf = nullptr;
  cudaMalloc(&buf, size_t(ROWS) * COLS * 2);
  cudaMemset(buf, 0, size_t(ROWS) * COLS * 2);
  long long* cyc = nullptr;
  cudaMalloc(&cyc, 256 * sizeof(long long));
  const uint64_t dims[2] = {COLS, ROWS};
  const uint64_t strides[1] = {uint64_t(COLS) * 2};
  CUtensorMap map_ld, map_st64, map_st128;
  const uint32_t box_ld[2] = {64, 128}, box64[2] = {32, 32}, box128[2] = {64, 32};
  if (make_tmap_bf16(&map_ld, buf, 2, dims, strides, box_ld, CU_TENSOR_MAP_SWIZZLE_128B) ||
      make_tmap_bf16(&map_st64, buf, 2, dims, strides, box64, CU_TENSOR_MAP_SWIZZLE_64B) ||
      make_tmap_bf16(&map_st128, buf, 2, dims, strides, box128, CU_TENSOR_MAP_SWIZZLE_128B)) {
    printf("tensor map: %s\n", mvd_last_error());
    return 1;
  }
  cudaFuncSetAttribute(load_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  cudaFuncSetAttribute(store_bw_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  cudaFuncSetAttribute(store_bw_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 2000;


  if (argc > 1 && argv[1][0] == '-' && argv[1][1] == '-' && argv[1][2] == 'd') {  // --depth: ring depth x sharing
    cudaFuncSetAttribute(load_depth_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 2048);
    cudaFuncSetAttribute(load_depth_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 2048);
    cudaFuncSetAttribute(load_depth_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 16384 + 2048);
    std::vector<long long> hh(256);
    for (int shared_rows : {0, 4096, 256}) {  // private boxes / a 10 MB window / a 0.6 MB window shared by all CTAs
      for (int depth : {4, 8, 12}) {
        for (int rep = 0; rep < 2; ++rep) {
          if (depth == 4) load_depth_kernel<4><<<148, 128, 4 * 16384 + 2048>>>(map_ld, iters, shared_rows, cyc);
          if (depth == 8) load_depth_kernel<8><<<148, 128, 8 * 16384 + 2048>>>(map_ld, iters, shared_rows, cyc);
          if (depth == 12) load_depth_kernel<12><<<148, 128, 12 * 16384 + 2048>>>(map_ld, iters, shared_rows, cyc);
          cudaDeviceSynchronize();
        }
        cudaMemcpy(hh.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = hh[i] > mx ? hh[i] : mx;
        printf("loads, 148 CTAs, %2d x 16 KB in flight, %-22s: %6.2f B/clk/SM\n", depth,
               shared_rows == 0 ? "private boxes" : (shared_rows == 4096 ? "shared 10 MB window" : "shared 0.6 MB window"),
               double(iters) * 16384 / double(mx));
      }
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
  }
  if (argc > 1) {  // --mix: loads and stores together at 148 CTAs
    cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    long long* cyc2 = nullptr;
    cudaMalloc(&cyc2, 512 * sizeof(long long));
    std::vector<long long> hh(512);
    const char* names[4] = {"loads only", "loads + TMA stores", "loads + st.global.v4", "st.global.v4 only"};
    for (int mode = 0; mode < 4; ++mode) {
      cudaMemset(cyc2, 0, 512 * sizeof(long long));
      for (int rep = 0; rep < 2; ++rep)
        mix_kernel<<<148, 160, 100 * 1024>>>(map_ld, map_st64, static_cast<__nv_bfloat16*>(buf), iters, mode, cyc2);
      cudaDeviceSynchronize();
      cudaMemcpy(hh.data(), cyc2, 512 * sizeof(long long), cudaMemcpyDeviceToHost);
      long long ml = 0, ms = 0;
      for (int i = 0; i < 148; ++i) {
        ml = hh[i] > ml ? hh[i] : ml;
        ms = hh[256 + i] > ms ? hh[256 + i] : ms;
      }
      printf("%-24s load %6.2f B/clk/SM (%lld clk)   store %6.2f B/clk/SM (%lld clk)\n", names[mode],
             ml ? double(iters) * 16384 / double(ml) : 0.0, ml, ms ? double(iters) * 4 * 2048 / double(ms) : 0.0, ms);
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
  }
  std::vector<long long> h(256);
  auto report = [&](const char* name, int ctas, double bytes_per_cta, float ms) {
    cudaMemcpy(h.data(), cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-28s CTAs %3d: %8.1f GB/s total, %6.1f GB/s/SM, %6.2f B/clk/SM (max %lld clk, %.3f ms)\n", name, ctas,
           bytes_per_cta * ctas / ms / 1e6, bytes_per_cta / ms / 1e6, bytes_per_cta / double(mx), mx, ms);
  };
  for (int ctas : {1, 37, 74, 148}) {
    for (int rep = 0; rep < 2; ++rep) {  // second repetition: L2-warm
      cudaEventRecord(e0);
      load_bw_kernel<<<ctas, 128, 66 * 1024>>>(map_ld, iters, cyc);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    report("load 128x128B (SW128)", ctas, double(iters) * 16384, time_ms(e0, e1));
  }
  for (int ctas : {1, 37, 74, 148}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      store_bw_kernel<64><<<ctas, 128, 66 * 1024>>>(map_st64, iters, cyc);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    report("store 4 warps x 32x64B (SW64)", ctas, double(iters) * 4 * 2048, time_ms(e0, e1));
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      store_bw_kernel<128><<<ctas, 128, 66 * 1024>>>(map_st128, iters, cyc);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    report("store 4 warps x 32x128B (SW128)", ctas, double(iters) * 4 * 4096, time_ms(e0, e1));
  }
  const cudaError_t err = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(err));
  return err != cudaSuccess;
}
