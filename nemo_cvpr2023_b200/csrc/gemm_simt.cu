// Generic fp32 GEMM with fused epilogue on the CUDA cores.  It serves the shapes the tcgen05 3xTF32
// kernel (gemm_tc.cu) does not take (tiny M/N/K, unaligned leading dimensions) and is the in-library
// cross-check for it in the tests.  128x128x16 tiles, 256 threads, 8x8 register micro-tile.
#include "common.cuh"

using namespace nb;

namespace {
constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_gemm_simt(const float* __restrict__ A, const float* __restrict__ B,
                                                   const float* __restrict__ bias, float* __restrict__ C, int M, int N,
                                                   int K, int lda, int ldb, int ldc, int act,
                                                   const float* __restrict__ gate, int ldg, float gslope, int accum) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int bm = blockIdx.y * BM, bn = blockIdx.x * BN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile -> As[k][m]
    if (!TA) {  // k contiguous in memory
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        int m = (tid >> 4) + 16 * p, k = tid & 15;
        int gm = bm + m, gk = k0 + k;
        As[k][m] = (gm < M && gk < K) ? A[(size_t)gm * lda + gk] : 0.f;
      }
    } else {  // m contiguous
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        int m = tid & 127, k = (tid >> 7) + 2 * p;
        int gm = bm + m, gk = k0 + k;
        As[k][m] = (gm < M && gk < K) ? A[(size_t)gk * lda + gm] : 0.f;
      }
    }
    if (!TB) {  // n contiguous
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        int n = tid & 127, k = (tid >> 7) + 2 * p;
        int gn = bn + n, gk = k0 + k;
        Bs[k][n] = (gn < N && gk < K) ? B[(size_t)gk * ldb + gn] : 0.f;
      }
    } else {  // k contiguous
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        int n = (tid >> 4) + 16 * p, k = tid & 15;
        int gn = bn + n, gk = k0 + k;
        Bs[k][n] = (gn < N && gk < K) ? B[(size_t)gn * ldb + gk] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = bm + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = bn + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (act == 1) v = fmaxf(v, 0.f);
      else if (act == 2) v = v > 0.f ? v : 0.01f * v;
      if (gate) v *= (gate[(size_t)m * ldg + n] > 0.f ? 1.f : gslope);
      float* c = C + (size_t)m * ldc + n;
      *c = accum ? *c + v : v;
    }
  }
}

// column sums: out[n] = sum_m X[m][n]; grid.x over 32-column strips, grid.y over row chunks (atomics)
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ X, float* __restrict__ out, int M, int N,
                                                int ld, int rows_per_block) {
  __shared__ float red[8][33];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int n = blockIdx.x * 32 + lane;
  int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int m = m0 + w; m < m1; m += 8) s += X[(size_t)m * ld + n];
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][lane];
    atomicAdd(out + n, t);
  }
}
}  // namespace

namespace nb {
int gemm_simt(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda, int ldb,
              int ldc, int transA, int transB, int act, const float* gate, int ldg, float gslope, int accumulate,
              cudaStream_t s) {
  dim3 grid(cdiv(N, BN), cdiv(M, BM));
#define NB_G(TA, TB) k_gemm_simt<TA, TB><<<grid, 256, 0, s>>>(A, B, bias, C, M, N, K, lda, ldb, ldc, act, gate, ldg, gslope, accumulate)
  if (transA) { if (transB) NB_G(true, true); else NB_G(true, false); }
  else { if (transB) NB_G(false, true); else NB_G(false, false); }
#undef NB_G
  NB_LAUNCH_CHECK("gemm_simt");
}
}  // namespace nb

extern "C" int nemo_b200_gemm_simt(const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
                                   int lda, int ldb, int ldc, int transA, int transB, int act, const float* gate,
                                   int ldg, float gate_slope, int accumulate, void* stream) {
  if (M == 0 || N == 0) return 0;
  NB_REQUIRE(A && B && C && M >= 0 && N >= 0 && K >= 0 && act >= 0 && act <= 2, "gemm: bad args");
  return nb::gemm_simt(A, B, bias, C, M, N, K, lda, ldb, ldc, transA, transB, act, gate, ldg, gate_slope, accumulate,
                       S(stream));
}

extern "C" int nemo_b200_colsum(const float* X, float* out, int M, int N, int ld, int accumulate, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(X && out && M >= 0 && N >= 0 && ld >= N, "colsum: bad args");
  if (!accumulate) NB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, S(stream)));
  if (M == 0) return 0;
  int rows = 256;
  dim3 grid(cdiv(N, 32), cdiv(M, rows));
  k_colsum<<<grid, 256, 0, S(stream)>>>(X, out, M, N, ld, rows);
  NB_LAUNCH_CHECK("colsum");
}
