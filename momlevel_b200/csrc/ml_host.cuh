// ml_host.cuh -- host-side plumbing shared by the translation units of libmomlevel_b200:
// thread-local error text, argument checks, launch accounting.
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/momlevel_b200.h"

namespace ml {

struct ThreadState {
  char err[512];
  int last_path;
  int force_direct;
  int64_t launches;
  int variants_chunk;  // main chunk width of the one-pass three-height kernel (0 = built-in default)
  const double* p_col;  // ml_set_column_pressure: per-column pressure offset (device, [ncol]) or NULL
  int64_t p_col_n;
};
ThreadState& tls();

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls().err, sizeof(tls().err), fmt, ap);
  va_end(ap);
  return code;
}

inline int cuda_fail(cudaError_t e, const char* what) {
  snprintf(tls().err, sizeof(tls().err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

// every kernel launch goes through this so gpu_launches in bench.py is a count, not a guess
inline int launched(const char* what) {
  tls().launches++;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ML_OK : cuda_fail(e, what);
}

#define ML_REQUIRE_PTR(p)                                             \
  do {                                                                \
    if ((p) == nullptr) return ml::fail(ML_ERR_NULL, "%s is NULL", #p); \
  } while (0)

#define ML_REQUIRE_ALIGNED(p, bytes)                                                        \
  do {                                                                                      \
    if ((reinterpret_cast<uintptr_t>(p) % (bytes)) != 0)                                    \
      return ml::fail(ML_ERR_ALIGN, "%s is not %d-byte aligned", #p, (int)(bytes));          \
  } while (0)

#define ML_CUDA(call)                                       \
  do {                                                      \
    cudaError_t e__ = (call);                               \
    if (e__ != cudaSuccess) return ml::cuda_fail(e__, #call); \
  } while (0)

inline int elem_size(int dtype) { return dtype == ML_F32 ? 4 : 8; }

}  // namespace ml
