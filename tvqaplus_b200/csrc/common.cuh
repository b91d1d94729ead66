// Host-side helpers shared by the C-ABI entry points: thread-local error text, argument checks, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/stage_b200.h"

namespace stg {

inline char* err_buf() {
  static thread_local char buf[512] = "";
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define STG_REQUIRE(cond, code, ...)                 \
  do {                                               \
    if (!(cond)) return ::stg::fail(code, __VA_ARGS__); \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(STAGE_E_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
  return STAGE_OK;
}

template <typename K>
inline int set_smem(K kernel, size_t bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return fail(STAGE_E_LAUNCH, "%s: cudaFuncSetAttribute(%zu B): %s", what, bytes, cudaGetErrorString(e));
  return STAGE_OK;
}

inline int num_sms() {  // of the CURRENT device (a process may drive several)
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (sms[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = v > 0 ? v : 148;
  }
  return sms[dev];
}

}  // namespace stg
