// Shared host/device helpers for librcmvs (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rcmvs.h"

namespace rcmvs {

void set_error(const char* fmt, ...);

inline cudaStream_t as_stream(rcmvs_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define RCMVS_REQUIRE(cond, code, ...)          \
  do {                                          \
    if (!(cond)) {                              \
      ::rcmvs::set_error(__VA_ARGS__);          \
      return (code);                            \
    }                                           \
  } while (0)

#define RCMVS_CHECK_LAUNCH(name)                                                         \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      ::rcmvs::set_error("%s: launch failed: %s", (name), cudaGetErrorString(e__));      \
      return RCMVS_ERR_CUDA;                                                             \
    }                                                                                    \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-(function, device) property: remember it per device ordinal
// (a process-wide flag would leave every GPU but the first at the 48 KB default -> "invalid argument" at launch).
struct SmemAttrOnce {
  unsigned long long done[4] = {0, 0, 0, 0};        // one bit per device ordinal (< 256)
  template <typename K>
  cudaError_t ensure(K kern, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    unsigned long long& word = done[(dev >> 6) & 3];
    if (__atomic_load_n(&word, __ATOMIC_ACQUIRE) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) __atomic_fetch_or(&word, bit, __ATOMIC_RELEASE);
    return e;
  }
};

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
// streaming store: written once, consumed by a later kernel from L2/HBM — do not pollute L1
__device__ __forceinline__ void st4_stream(float* p, const float4& v) { __stcs(reinterpret_cast<float4*>(p), v); }

}  // namespace rcmvs
