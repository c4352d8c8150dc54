// nemo_b200_gemm: shape dispatch between the tcgen05 3xTF32 kernels (gemm_tc.cu: 128 x 128 x 32 tiles; gemm_tc2.cu:
// 128 x 256 x 16 tiles, N = 256 MMAs) and the CUDA-core kernel (gemm_simt.cu).  All are kernels of this library; there is
// no library (cuBLAS) or CPU path.
#include <stdlib.h>
#include "common.cuh"

namespace nb {
int gemm_simt(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda, int ldb,
              int ldc, int transA, int transB, int act, const float* gate, int ldg, float gslope, int accumulate,
              cudaStream_t s);
bool gemm_tc_applicable(const float* A, const float* B, int M, int N, int K, int lda, int ldb);
int gemm_tc(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda, int ldb,
            int ldc, int transA, int transB, int act, const float* gate, int ldg, float gslope, int accumulate,
            int force_splits, cudaStream_t s);
double gemm_tc2_cost(int M, int N, int K, int sms, int* splits_out);
bool gemm_tc2_applicable(const float* C, int N, int ldc, const float* gate, int ldg);
int gemm_tc2(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda, int ldb,
             int ldc, int transA, int transB, int act, const float* gate, int ldg, float gslope, int accumulate,
             int force_splits, cudaStream_t s);

// estimated tensor-pipe cycles of the busiest SM for the 128 x 128 x 32 kernel (12 MMAs of ~96 cycles per k-block)
static double gemm_tc1_cost(int M, int N, int K, int act, int sms) {
  const int tiles = cdiv(M, 128) * cdiv(N, 128), kblocks = cdiv(K, 32);
  int splits = 1;
  if (act == 0 && tiles < sms) splits = max(1, min(sms / tiles, kblocks / 4));
  return (double)cdiv(tiles * splits, sms) * (cdiv(kblocks, splits) * 1190.0 + 4000.0);
}
// NEMO_B200_GEMM_KERNEL=1 / 2 forces a generation (experiments); default: the cheaper estimate
static bool prefer_tc2(const float* C, int ldc, const float* gate, int ldg, int M, int N, int K, int act) {
  static const int forced = [] { const char* e = getenv("NEMO_B200_GEMM_KERNEL"); return e ? atoi(e) : 0; }();
  if (!gemm_tc2_applicable(C, N, ldc, gate, ldg) || forced == 1) return false;
  if (forced == 2) return true;
  int sp;
  return gemm_tc2_cost(M, N, K, num_sms(), act == 0 ? &sp : nullptr) < gemm_tc1_cost(M, N, K, act, num_sms());
}
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
  if (nb::gemm_tc_applicable(A, B, M, N, K, lda, ldb)) {
    if (nb::prefer_tc2(C, ldc, gate, ldg, M, N, K, act))
      return nb::gemm_tc2(A, B, bias, C, M, N, K, lda, ldb, ldc, transA, transB, act, gate, ldg, gate_slope, accumulate,
                          0, nb::S(stream));
    return nb::gemm_tc(A, B, bias, C, M, N, K, lda, ldb, ldc, transA, transB, act, gate, ldg, gate_slope, accumulate,
                       0, nb::S(stream));
  }
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

extern "C" int nemo_b200_gemm_tc2(const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
                                  int lda, int ldb, int ldc, int transA, int transB, int act, const float* gate,
                                  int ldg, float gate_slope, int accumulate, int splits, void* stream) {
  if (M == 0 || N == 0) return 0;
  NB_GEMM_CHECK();
  NB_REQUIRE(K >= 1 && ((((uintptr_t)A | (uintptr_t)B) & 15) == 0) && ((lda | ldb) & 3) == 0,
             "gemm_tc2: operands must be 16-byte aligned with leading dimensions that are multiples of 4");
  NB_REQUIRE(nb::gemm_tc2_applicable(C, N, ldc, gate, ldg),
             "gemm_tc2: C and gate must be 16-byte addressable (ld % 4 == 0; N % 4 == 0 or ldc - N < 4)");
  return nb::gemm_tc2(A, B, bias, C, M, N, K, lda, ldb, ldc, transA, transB, act, gate, ldg, gate_slope, accumulate,
                      splits, nb::S(stream));
}
