// Host-side helpers shared by the C-ABI entry points: error reporting and TMA tensor-map encoding
// (cuTensorMapEncodeTiled resolved through the runtime so the library does not link libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <atomic>

namespace tsnet {

inline char* last_error_buf() {
  static thread_local char buf[512] = "";
  return buf;
}
inline int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
#define TSNET_ARG_CHECK(cond, ...) \
  do {                             \
    if (!(cond)) return ::tsnet::set_error(-1, __VA_ARGS__); \
  } while (0)
#define TSNET_CUDA_CHECK(expr)                                                                  \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::tsnet::set_error(static_cast<int>(_e), "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

// every kernel launch of the library goes through this check: it also counts the launch (tsnet_launch_count(),
// reported by bench.py as gpu_launches)
std::atomic<long long>& launch_counter();
#define TSNET_LAUNCH_CHECK()            \
  do {                                  \
    ++::tsnet::launch_counter();        \
    TSNET_CUDA_CHECK(cudaGetLastError()); \
  } while (0)

// Per-device state: function attributes (cudaFuncSetAttribute) and the SM count belong to a device / context, not to
// the process, so the "done once" flags are kept per device ordinal.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < kMaxDevices ? dev : 0;
}
// opt a kernel in to `bytes` of dynamic shared memory on the current device (once per device and size)
template <typename K>
inline cudaError_t ensure_dyn_smem(K kernel, int bytes, int (&cache)[kMaxDevices]) {
  const int dev = current_device();
  if (cache[dev] >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) cache[dev] = bytes;
  return e;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 16-bit element tensor map, 128 B swizzle, innermost box = 64 elements (128 B).
// dims / box innermost first; strides in BYTES for dims 1..rank-1.
inline int encode_tmap_u16_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                                 const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_encodeTiled fn = get_encode_tiled();
  if (!fn) return set_error(-2, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-3, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

inline int num_sms() {
  static int n[kMaxDevices] = {0};
  const int dev = current_device();
  if (!n[dev]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

}  // namespace tsnet
