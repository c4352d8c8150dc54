// Shared helpers for libnemo_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/nemo_b200.h"

namespace nb {

void set_error(const char* fmt, ...);

#define NB_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) {                                              \
      nb::set_error(__VA_ARGS__);                               \
      return NEMO_B200_EINVAL;                                  \
    }                                                           \
  } while (0)

#define NB_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      nb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return NEMO_B200_ECUDA;                                                           \
    }                                                                                   \
  } while (0)

// after a kernel launch
#define NB_LAUNCH_CHECK(name)                                                           \
  do {                                                                                  \
    cudaError_t e__ = cudaPeekAtLastError();                                            \
    if (e__ != cudaSuccess) {                                                           \
      (void)cudaGetLastError();                                                         \
      nb::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));          \
      return NEMO_B200_ECUDA;                                                           \
    }                                                                                   \
    return 0;                                                                           \
  } while (0)

#define NB_LAUNCH_CHECK_CONT(name)                                                      \
  do {                                                                                  \
    cudaError_t e__ = cudaPeekAtLastError();                                            \
    if (e__ != cudaSuccess) {                                                           \
      (void)cudaGetLastError();                                                         \
      nb::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));          \
      return NEMO_B200_ECUDA;                                                           \
    }                                                                                   \
  } while (0)

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int NV = NEMO_B200_NV;
constexpr int VP = NEMO_B200_VP;
constexpr int NJ = NEMO_B200_NJ;

#define NB_HD __host__ __device__ __forceinline__

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace nb
