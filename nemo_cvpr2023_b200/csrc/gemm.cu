// nemo_b200_gemm: shape dispatch between the tcgen05 3xTF32 kernel and the CUDA-core kernel.
#include "common.cuh"

namespace nb {
int gemm_simt(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda, int ldb,
              int ldc, int transA, int transB, int act, const float* gate, int ldg, float gslope, int accumulate,
              cudaStream_t s);
}

extern "C" int nemo_b200_gemm(const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
                              int lda, int ldb, int ldc, int transA, int transB, int act, const float* gate, int ldg,
                              float gate_slope, int accumulate, void* stream) {
  if (M == 0 || N == 0) return 0;
  NB_REQUIRE(A && B && C && M >= 0 && N >= 0 && K >= 0 && act >= 0 && act <= 2, "gemm: bad args");
  NB_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "gemm: leading dimension too small");
  return nb::gemm_simt(A, B, bias, C, M, N, K, lda, ldb, ldc, transA, transB, act, gate, ldg, gate_slope, accumulate,
                       nb::S(stream));
}
