// host_common.h — host-side helpers shared by the launchers: error reporting, TMA tensor-map creation
// through the driver entry point (no link-time dependency on libcuda, so the library loads on a
// CPU-only box for the symbol checks), SM count cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace mvd {

void set_error(const char* fmt, ...);
int sm_count();
void count_launches(int n);  // bookkeeping behind mvd_kernel_launch_count()

// Returns 0 on success. dims/strides innermost-first; strides in BYTES for dims 1..rank-1.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, CUtensorMapSwizzle swizzle);

// Programmatic dependent launch (PDL): the kernel may be scheduled while its stream predecessor is still draining;
// every kernel launched this way executes pdl_wait() before it touches global memory (see below), so ordering and
// visibility are exactly those of a normal stream launch, but launch latency and prologues (barrier init, TMEM
// allocation, descriptor prefetch) overlap the predecessor's tail. Opt-in with MVD_PDL=1 (round 1 measured no gain
// for the graph-replayed step, see host_common.cu).
bool pdl_enabled();
void set_pdl(int on);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
// all threads, before the first global-memory access that may depend on the previous kernel in the stream
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// allow the next kernel in the stream to be scheduled (it still waits for our completion in its own pdl_wait())
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

#define MVD_CHECK(cond, ...)          \
  do {                                \
    if (!(cond)) {                    \
      mvd::set_error(__VA_ARGS__);    \
      return MVD_ERR_INVALID;         \
    }                                 \
  } while (0)

#define MVD_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) {                                                  \
      mvd::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));         \
      return MVD_ERR_CUDA;                                                    \
    }                                                                         \
  } while (0)

}  // namespace mvd
