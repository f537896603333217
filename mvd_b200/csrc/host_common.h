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
