// nemo_b200_gemm: shape dispatch between the tcgen05 3xTF32 kernel (gemm_tc.cu) and the CUDA-core kernel
// (gemm_simt.cu).  Both are kernels of this library; there is no library (cuBLAS) or CPU path.
#include "common.cuh"

namespace nb {
int gemm_simt(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda, int ldb,
              int ldc, int transA, int transB, int act, const float* gate, int ldg, float gslope, int accumulate,
              cudaStream_t s);
bool gemm_tc_applicable(const float* A, const float* B, int M, int N, int K, int lda, int ldb);
int gemm_tc(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda, int ldb,
            int ldc, int transA, int transB, int act, const float* gate, int ldg, float gslope, int accumulate,
            int force_splits, cudaStream_t s);
}  // namespace nb

#define NB_GEMM_CHECK()                                                                                           \
  NB_REQUIRE(A && B && C && M >= 0 && N >= 0 && K >= 0 && act >= 0 && act <= 2, "gemm: bad args");                 \
  NB_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "gemm: leading dimension too small"); \
  NB_REQUIRE(!gate || ldg >= N, "gemm: gate leading dimension too small")

extern "C" int nemo_b200_gemm(const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
                              int lda, int ldb, int ldc, int transA, int transB, int act, const float* gate, int ldg,
                              float gate_slope, int accumulate, void* stream) {
  if (M == 0 || N == 0) return 0;
  NB_GEMM_CHECK();
  if (nb::gemm_tc_applicable(A, B, M, N, K, lda, ldb))
    return nb::gemm_tc(A, B, bias, C, M, N, K, lda, ldb, ldc, transA, transB, act, gate, ldg, gate_slope, accumulate,
                       0, nb::S(stream));
  return nb::gemm_simt(A, B, bias, C, M, N, K, lda, ldb, ldc, transA, transB, act, gate, ldg, gate_slope, accumulate,
                       nb::S(stream));
}

extern "C" int nemo_b200_gemm_tc(const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
                                 int lda, int ldb, int ldc, int transA, int transB, int act, const float* gate,
                                 int ldg, float gate_slope, int accumulate, int splits, void* stream) {
  if (M == 0 || N == 0) return 0;
  NB_GEMM_CHECK();
  NB_REQUIRE(K >= 1 && ((((uintptr_t)A | (uintptr_t)B) & 15) == 0) && ((lda | ldb) & 3) == 0,
             "gemm_tc: operands must be 16-byte aligned with leading dimensions that are multiples of 4");
  return nb::gemm_tc(A, B, bias, C, M, N, K, lda, ldb, ldc, transA, transB, act, gate, ldg, gate_slope, accumulate,
                     splits, nb::S(stream));
}
