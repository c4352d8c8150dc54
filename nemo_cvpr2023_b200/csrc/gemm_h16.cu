// K5 forward on fp16 operand PAIRS: C[M,N] = out_scale . (a_scale A)[M,K] . (Bh + Bl)[K,N]
// Reference arithmetic: the pose-blend matmul of lbs (human_body_prior/body_model/lbs.py:229-233),
// pose_offsets = (R[1:] - I).view(N, 207) @ posedirs.
//
// A tcgen05.mma.kind::f16 costs the same cycles for K = 16 as a kind::tf32 one for K = 8 (tools/ubench/mma_f16.cu) and
// an fp16 pair hi = rn16(x), lo = rn16(x - hi) carries the same 22 mantissa bits as a tf32 pair: the 3-product scheme
// A.B ~= A_hi.B_hi + A_hi.B_lo + A_lo.B_hi needs HALF the MMAs of 3xTF32.  fp16 has 5 exponent bits, so the operands
// must be brought into range by a power of two -- which is why this kernel exists for ONE GEMM only, the one whose
// operand ranges are known: the pose feature R - I lies in [-2, 2] (a_scale = 2^13), posedirs is a constant of the body
// model (split once on the host into Bh / Bl (N, ldb) K-major with its own power of two).  Elements far below the
// largest keep an ABSOLUTE precision of 2^-25 of the scaled range (fp16 subnormals), i.e. 2^-38 of the largest
// product: far below the fp32 rounding of the sum.  out_scale = 1 / (a_scale b_scale) is exact.
//
// Structure = the second-generation 3xTF32 kernel (gemm_tc2.cu) with a stage of 32 k: TMA brings the raw fp32 A tile
// (128 rows x 128 bytes, 128-byte swizzle) and the two fp16 B tiles (256 rows x 64 bytes, 64-byte swizzle -- no B
// converter warps); the A converter warps scale, split and pack the tile into tensor memory (two halves per column:
// TS-mode MMAs); 6 MMAs of N = 256 per stage (828 cycles per 32 k, against 1 656 for 3xTF32); one 256-column accumulator
// folded into fp32 registers; epilogue through swizzled 32 x 32 boxes and coalesced 16-byte stores.
#include <stdlib.h>
#include <cuda.h>
#include <cuda_fp16.h>
#include <map>
#include <mutex>
#include <array>
#ifndef NB_MBAR_SUSPEND_NS
#define NB_MBAR_SUSPEND_NS 1000000
#endif
#include "common.cuh"
#include "sm100.cuh"

using namespace nb;
using namespace nb::sm100;

namespace {

constexpr int BM = 128, BN = 256, BKS = 32;        // tile; k per stage
constexpr int NUM_THREADS = 512;
constexpr int A_BYTES = BM * BKS * 4;              // 16 KB raw fp32
constexpr int BH_BYTES = BN * BKS * 2;             // 16 KB per fp16 part
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = A_BYTES + 2 * BH_BYTES;  // [A raw][B hi][B lo] = what TMA delivers
constexpr int BOX_BYTES = 32 * 32 * 4;
constexpr int OFF_STAGING = STAGES * STAGE_BYTES;  // 8 warps x 2 boxes
constexpr int STAGING_BYTES = 8 * 2 * BOX_BYTES;
constexpr int OFF_BARS = OFF_STAGING + STAGING_BYTES;
constexpr int SMEM_BYTES = OFF_BARS + 256 + 1024;
constexpr int TMEM_COLS = 512, TMEM_A = BN;        // accumulator 256 | A slots 3 x (16 hi + 16 lo) columns
constexpr int NC = BN / 2;                         // accumulator columns per epilogue thread
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(STAGE_BYTES % 1024 == 0 && A_BYTES % 1024 == 0 && BH_BYTES % 1024 == 0, "swizzled operand alignment");

struct Params {
  int M, N, K, mtiles, ntiles, kblocks, flush, splits;
  float a_scale, out_scale;
  const float* bias;
  float* C;
  int ldc;
};

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(NUM_THREADS, 1) k_gemm_h16(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmBh,
                                                             const __grid_constant__ CUtensorMap tmBl,
                                                             const __grid_constant__ CUtensorMap tmC, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + OFF_BARS;
  auto full = [&](int s) { return bars + 8u * s; };                  // TMA landed                      (tx)
  auto conv = [&](int s) { return bars + 8u * (STAGES + s); };       // A pair in TMEM                   (128)
  auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };  // MMAs done with the stage         (commit)
  const uint32_t pfull = bars + 8u * (3 * STAGES);                   // partial accumulator complete     (commit)
  const uint32_t pempty = bars + 8u * (3 * STAGES + 1);              // partial folded into registers    (256)
  const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmBh);
    tma_prefetch_desc(&tmBl);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(conv(s), 128);
      mbar_init(empty(s), 1);
    }
    mbar_init(pfull, 1);
    mbar_init(pempty, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32u(tmem_slot);
  // unit = (tile, k split): a long K (the dgrad, K = 20 736) is cut so that every SM has work; the partial tiles then
  // leave through TMA reduce-add into the zeroed C
  const int units = p.mtiles * p.ntiles * p.splits;
  constexpr uint32_t idesc = idesc_f16(BM, BN);
  auto krange = [&](int split, int& kb0, int& kb1) {
    kb0 = (int)((long long)split * p.kblocks / p.splits);
    kb1 = (int)((long long)(split + 1) * p.kblocks / p.splits);
  };

  if (warp < 4) {
    setmaxnreg_dec<40>();
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer
      if (lane == 0) {
        uint32_t it = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
          const int split = u % p.splits, tile = u / p.splits;
          const int mb = tile % p.mtiles, nb = tile / p.mtiles;
          int kb0, kb1;
          krange(split, kb0, kb1);
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const int s = it % STAGES;
            mbar_wait(empty(s), ((it / STAGES) & 1) ^ 1);
            mbar_arrive_expect_tx(full(s), STAGE_BYTES);
            const uint32_t sa = sbase + s * STAGE_BYTES;
            tma_load_2d(sa, &tmA, full(s), kb * BKS, mb * BM);
            tma_load_2d(sa + A_BYTES, &tmBh, full(s), kb * BKS, nb * BN);
            tma_load_2d(sa + A_BYTES + BH_BYTES, &tmBl, full(s), kb * BKS, nb * BN);
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer
      if (lane == 0) {
        uint32_t it = 0, pc = 0;
        const uint32_t d = tmem_base;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
          int kb0, kb1;
          krange(u % p.splits, kb0, kb1);
          for (int kc = kb0; kc < kb1; kc += p.flush, ++pc) {
            if (pc > 0) mbar_wait(pempty, (pc - 1) & 1);   // the previous partial has been folded into registers
            tc_fence_after();
            const int kce = min(kc + p.flush, kb1);
            for (int kb = kc; kb < kce; ++kb, ++it) {
              const int s = it % STAGES;
              mbar_wait(conv(s), (it / STAGES) & 1);
              tc_fence_after();
              const uint32_t a_hi = tmem_base + TMEM_A + s * 32, a_lo = a_hi + 16;
              const uint32_t b_hi = sbase + s * STAGE_BYTES + A_BYTES, b_lo = b_hi + BH_BYTES;
              // K-major fp16: 64-byte rows (32 k), 8-row groups of 512 bytes, 64-byte swizzle; a k step of 16 = 32 bytes
              const uint64_t bh0 = umma_smem_desc(b_hi, 16, 512, LAYOUT_SW64);
              const uint64_t bl0 = umma_smem_desc(b_lo, 16, 512, LAYOUT_SW64);
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint64_t bh = umma_desc_advance(bh0, j * 32), bl = umma_desc_advance(bl0, j * 32);
                umma_f16_ts(d, a_lo + j * 8, bh, idesc, (kb > kc || j > 0) ? 1u : 0u);
                umma_f16_ts(d, a_hi + j * 8, bl, idesc, 1u);
                umma_f16_ts(d, a_hi + j * 8, bh, idesc, 1u);
              }
              umma_commit(empty(s));  // stage (raw tiles, TMEM A slot) reusable once these MMAs have completed
            }
            umma_commit(pfull);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ A converters (128 threads, thread = row)
    setmaxnreg_dec<72>();
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_A;
    const float sc = p.a_scale;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int kb0, kb1;
      krange(u % p.splits, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t sa = sbase + s * STAGE_BYTES;
        mbar_wait(full(s), (it / STAGES) & 1);
        // raw tile: row r = 128 bytes (32 k), 16-byte chunk c stored at chunk c ^ (r & 7)
        uint32_t h[16], w[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 x = lds128(sa + row * 128 + ((c ^ (row & 7)) << 4));
          split_h2(x.x * sc, x.y * sc, h[2 * c], w[2 * c]);
          split_h2(x.z * sc, x.w * sc, h[2 * c + 1], w[2 * c + 1]);
        }
        tmem_st16(trow + s * 32, h);
        tmem_st16(trow + s * 32 + 16, w);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(conv(s));
      }
    }
  } else {
    // ------------------------------------------------------------------ accumulate + epilogue (256 threads)
    setmaxnreg_inc<184>();
    const int q = warp & 3, hc = (warp - 8) >> 2;  // TMEM lane quarter, column half
    const uint32_t box0 = sbase + OFF_STAGING + (warp - 8) * (2 * BOX_BYTES);
    const uint32_t rowo = lane * 128;
    uint32_t pc = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int split = u % p.splits, tile = u / p.splits;
      const int mb = tile % p.mtiles, nb = tile / p.mtiles;
      int kb0, kb1;
      krange(split, kb0, kb1);
      float acc[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) acc[j] = 0.f;
      for (int kc = kb0; kc < kb1; kc += p.flush, ++pc) {
        mbar_wait(pfull, pc & 1);
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + hc * NC;
#pragma unroll
        for (int c = 0; c < NC / 16; ++c) {
          uint32_t v[16];
          tmem_ld16(ta + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[c * 16 + j] += __uint_as_float(v[j]);
        }
        tc_fence_before();
        mbar_arrive(pempty);  // folded: the MMA warp may overwrite the accumulator
      }
      const int m_base = mb * BM + q * 32;
      const int ncol0 = nb * BN + hc * NC;
      if (m_base >= p.M || ncol0 >= p.N) continue;  // warp-uniform
      const float os = p.out_scale;
#pragma unroll
      for (int rd = 0; rd < 2; ++rd) {
        const int n_rd = ncol0 + rd * 64;
        if (n_rd >= p.N) break;  // warp-uniform
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int n0 = n_rd + 32 * b;
          if (n0 < p.N) {  // warp-uniform
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const int j0 = rd * 64 + b * 32 + 4 * c4;
              float x[4] = {acc[j0] * os, acc[j0 + 1] * os, acc[j0 + 2] * os, acc[j0 + 3] * os};
              if (p.bias && split == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) x[k] += (n0 + 4 * c4 + k < p.N) ? __ldg(p.bias + n0 + 4 * c4 + k) : 0.f;
              }
              sts128(box0 + b * BOX_BYTES + rowo + ((c4 ^ (lane & 7)) << 4), make_float4(x[0], x[1], x[2], x[3]));
            }
          }
        }
        if (p.splits > 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tmC, box0, n_rd, m_base);
            if (n_rd + 32 < p.N) tma_reduce_add_2d(&tmC, box0 + BOX_BYTES, n_rd + 32, m_base);
            bulk_commit();
            bulk_wait_read0();   // the boxes are reused by the next round / tile
          }
          __syncwarp();
          continue;
        }
        // read the boxes back row-wise (eight lanes per 128-byte row, four rows per instruction), 16-byte stores
        __syncwarp();
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int n0 = n_rd + 32 * b + 4 * (lane & 7);
          if (n_rd + 32 * b < p.N) {  // warp-uniform
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = 4 * i + (lane >> 3);
              const float4 v = lds128(box0 + b * BOX_BYTES + r * 128 + (((lane & 7) ^ (r & 7)) << 4));
              if (m_base + r < p.M && n0 < p.N)
                *reinterpret_cast<float4*>(p.C + (size_t)(m_base + r) * p.ldc + n0) = v;
            }
          }
        }
        __syncwarp();
      }
    }
  }
  if (warp >= 8 && lane == 0) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

// row-major matrix [outer][inner] of 4-byte (fp32) or 2-byte (fp16) elements, row stride ld elements
int make_map(CUtensorMap* out, const void* ptr, int elem, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
             uint32_t box_outer, CUtensorMapSwizzle sw) {
  static std::mutex mu;
  static std::map<std::array<uint64_t, 8>, CUtensorMap> cache;
  const std::array<uint64_t, 8> key = {(uint64_t)(uintptr_t)ptr, (uint64_t)elem, inner, outer, ld, box_inner, box_outer,
                                       (uint64_t)sw};
  std::lock_guard<std::mutex> lk(mu);
  auto f = cache.find(key);
  if (f != cache.end()) {
    *out = f->second;
    return 0;
  }
  EncodeTiledFn enc = encode_fn();
  NB_REQUIRE(enc, "gemm_h16: cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * (uint64_t)elem};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_h16: cuTensorMapEncodeTiled failed (%d) for %llu x %llu x %d B, ld %llu, box %u x %u", (int)r,
              (unsigned long long)outer, (unsigned long long)inner, elem, (unsigned long long)ld, box_inner, box_outer);
    return NEMO_B200_ECUDA;
  }
  if (cache.size() > 1024) cache.clear();
  cache[key] = *out;
  return 0;
}

}  // namespace

extern "C" int nemo_b200_gemm_h16(const float* A, const void* Bh, const void* Bl, const float* bias, float* C, int M, int N,
                                  int K, int lda, int ldb, int ldc, float a_scale, float out_scale, void* stream) {
  if (M == 0 || N == 0) return 0;
  NB_REQUIRE(A && Bh && Bl && C && M > 0 && N > 0 && K > 0, "gemm_h16: bad args");
  NB_REQUIRE(lda >= K && ldb >= K && ldc >= N, "gemm_h16: leading dimension too small");
  NB_REQUIRE((((uintptr_t)A | (uintptr_t)Bh | (uintptr_t)Bl | (uintptr_t)C) & 15) == 0 && (lda & 3) == 0 && (ldb & 7) == 0 &&
                 (ldc & 3) == 0 && ((N & 3) == 0 || ldc - N < 4),
             "gemm_h16: operands must be 16-byte aligned with 16-byte row strides");
  cudaStream_t s = S(stream);
  Params p;
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.C = C; p.ldc = ldc;
  p.a_scale = a_scale; p.out_scale = out_scale;
  p.mtiles = cdiv(M, BM);
  p.ntiles = cdiv(N, BN);
  p.kblocks = cdiv(K, BKS);
  p.flush = 8;   // k-blocks (of 32) accumulated in tensor memory before the fold into fp32 registers: 256 k values
  const int tiles = p.mtiles * p.ntiles, sms = num_sms();
  p.splits = 1;
  if (tiles < sms && !bias) p.splits = max(1, min(sms / tiles, p.kblocks / 8));
  if (p.splits > 1) NB_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, s));
  CUtensorMap ta, tbh, tbl, tc;
  if (int rc = make_map(&tc, C, 4, N, M, ldc, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = make_map(&ta, A, 4, K, M, lda, BKS, BM, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = make_map(&tbh, Bh, 2, K, N, ldb, BKS, BN, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  if (int rc = make_map(&tbl, Bl, 2, K, N, ldb, BKS, BN, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  const int grid = min(tiles * p.splits, sms);
  NB_SET_SMEM_ONCE(k_gemm_h16, SMEM_BYTES);
  k_gemm_h16<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta, tbh, tbl, tc, p);
  NB_LAUNCH_CHECK("gemm_h16");
}
