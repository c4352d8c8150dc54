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

// Per-device caches (an engine may live on any device of the process): SM count, and "this kernel's dynamic shared
// memory attribute has been set on this device".
constexpr int MAX_DEVICES = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < MAX_DEVICES ? dev : 0;
}
inline int num_sms() {
  static int n[MAX_DEVICES] = {0};
  const int dev = current_device();
  if (!n[dev]) {
    int v = 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v;
  }
  return n[dev];
}
// once per (kernel, device)
#define NB_SET_SMEM_ONCE(kernel, bytes)                                                              \
  do {                                                                                               \
    static bool done__[nb::MAX_DEVICES] = {false};                                                   \
    const int dev__ = nb::current_device();                                                          \
    if (!done__[dev__]) {                                                                            \
      NB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes)));   \
      done__[dev__] = true;                                                                          \
    }                                                                                                \
  } while (0)

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
// Reduce K (power of two <= 32) per-lane values across the warp; on return v[0] of lane l holds the
// total of value index (l & (K-1)) ... for K == 32 that is value l.  31 shuffles for K = 32.
template <int K>
__device__ __forceinline__ void warp_reduce_scatter(float* v, int lane) {
  int off = 16;
#pragma unroll
  for (int half = K / 2; half >= 1; half >>= 1, off >>= 1) {
    bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      float send = upper ? v[i] : v[i + half];
      float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}
// index of the value that lane ends up holding after warp_reduce_scatter<K>
template <int K>
__device__ __forceinline__ int scatter_index(int lane) {
  if (K == 32) return lane;
  if (K == 4) return ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
  return 0;
}

#endif

}  // namespace nb
