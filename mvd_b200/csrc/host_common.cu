// host_common.cu — error string, SM count, TMA tensor-map encoding via the driver entry point.
#include "host_common.h"
#include "../../include/mvd_b200.h"

#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>

namespace mvd {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Programmatic dependent launch. Measured on B200 under graph replay: with 8 samples per GPU the step is throughput
// bound and PDL costs 2-3 %; with 1 sample per GPU (view x CFG sharded rank) it gained 5 % mid-round (5.32 vs 5.61 ms)
// and loses 2 % on the final tree (stream-K / persistent kernels already cover every SM). So it is a run-time switch:
// MVD_PDL=0/1 forces it, otherwise the caller (DenoiseSession) sets it with mvd_set_launch_overlap() — off by default.
static std::atomic<int> g_pdl{-1};
static int pdl_env() {
  static const int v = [] {
    const char* e = getenv("MVD_PDL");
    return e == nullptr ? -1 : (e[0] == '1' ? 1 : 0);
  }();
  return v;
}
bool pdl_enabled() {
  const int forced = pdl_env();
  if (forced >= 0) return forced != 0;
  return g_pdl.load(std::memory_order_relaxed) > 0;
}
void set_pdl(int on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }

int sm_count() {  // of the CURRENT device (cached per device ordinal; 148 when no device is visible)
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return MVD_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                   gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
              rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
    return MVD_ERR_CUDA;
  }
  return MVD_OK;
}

}  // namespace mvd

extern "C" {
const char* mvd_last_error(void) { return mvd::g_err; }
int mvd_abi_version(void) { return MVD_ABI_VERSION; }
int64_t mvd_kernel_launch_count(void) { return static_cast<int64_t>(mvd::g_launches.load()); }
int mvd_set_launch_overlap(int on) {
  mvd::set_pdl(on);
  return MVD_OK;
}
}
