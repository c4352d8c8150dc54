// K2 / K5, second generation: the 3xTF32 tcgen05 GEMM with 128 x 256 tiles.
//
// A tcgen05.mma.kind::tf32 (M = 128, K = 8) costs ~96 issue cycles up to N ~ 176 and ~138 at N = 256
// (tools/ubench/mma_pace.cu): the 128 x 128 tiles of the first kernel run the tensor pipe at 72 % of what N = 256
// delivers, and the timeline of its CTA 0 (NEMO_B200_GEMM_TIMING) shows the main loop AT that bound (1 187 cycles per
// 128 x 128 x 32 block against 12 x 96).  This kernel issues N = 256:
//   * tile 128 x 256 x 16.  A k-block is 2 k-steps x 3 products = 6 MMAs (828 cycles); a stage is 40 KB -- A raw 8 KB,
//     B raw (= hi operand: the tensor core ignores the low 13 mantissa bits) 16 KB, B lo 16 KB -- so the ring stays
//     4 deep (a 32-wide k-block at N = 256 would leave room for two stages, less than one TMA + convert round trip);
//   * K-major operands are 64-byte rows (TMA / UMMA SWIZZLE_64B), MN-major ones 32 x 16 boxes with 32-byte atoms;
//   * ONE accumulator in tensor memory (256 columns) + four 32-column A slots (tf32 hi | lo, written by the converter
//     warps: TS-mode MMAs).  Every `flush` k-blocks the accumulate warps fold the partial into fp32 registers
//     (round-to-nearest; the tensor core accumulates with truncation) -- ~400 cycles during which the MMA warp waits,
//     3 % at flush = 16 -- so the register copy, not tensor memory, carries the tile into the epilogue and the next
//     tile's MMAs start right after the last fold;
//   * 128 accumulator registers per epilogue thread: the warpgroups re-divide the register file (setmaxnreg):
//     TMA / MMA / B-converter warps drop to 40, A converters to 72, the two accumulate warpgroups rise to 152 (the sum
//     stays within the 640 x 96 registers the CTA was launched with);
//   * epilogue as in the first kernel: bias / activation / gate in registers, 32 x 32 boxes through swizzled shared
//     memory, TMA store or TMA reduce-add (C +=, split-K); two boxes per warp in flight.
// 20 warps: 0 TMA producer, 1 MMA issuer, (2-3 idle), 4-7 A converters, 8-15 accumulate + epilogue, 16-19 B converters.
// nemo_b200_gemm picks this kernel or the 128 x 128 one by the estimated tensor-pipe time of the shape
// (waves x k-blocks x cycles per k-block): 4097 x 1000 x 1000 is one wave of 132 tiles here against two waves of 264.
#include <stdlib.h>
#include <cuda.h>
#include <map>
#include <mutex>
#include <array>
#include "common.cuh"
#ifndef NB_MBAR_SUSPEND_NS
#define NB_MBAR_SUSPEND_NS 1000000  // waiting warps are parked by the hardware instead of spinning on try_wait
#endif
#include "sm100.cuh"

using namespace nb;
using namespace nb::sm100;

namespace {

constexpr int BM = 128, BN = 256, BK = 16, UMMA_K = 8;
constexpr int NUM_THREADS = 640;
constexpr int A_BYTES = BM * BK * 4;               // 8 KB
constexpr int B_BYTES = BN * BK * 4;               // 16 KB
constexpr int STAGES = 4;
constexpr int RAW_BYTES = A_BYTES + B_BYTES;       // what TMA delivers per stage
constexpr int STAGE_BYTES = RAW_BYTES + B_BYTES;   // [A raw][B raw = hi][B lo]
constexpr int BOX_BYTES = 32 * 32 * 4;             // one 32 x 32 epilogue box
constexpr int OFF_STAGING = STAGES * STAGE_BYTES;  // 8 warps x 2 boxes
constexpr int STAGING_BYTES = 8 * 2 * BOX_BYTES;
constexpr int OFF_BARS = OFF_STAGING + STAGING_BYTES;
constexpr int SMEM_BYTES = OFF_BARS + 256 + 1024;
constexpr int TMEM_COLS = 512, TMEM_A = BN;        // accumulator 256 | A slots 4 x (16 hi + 16 lo)
constexpr int NC = BN / 2;                         // accumulator columns per epilogue thread
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(TMEM_A + STAGES * 2 * BK <= TMEM_COLS, "tensor memory budget");
static_assert(STAGE_BYTES % 1024 == 0 && A_BYTES % 1024 == 0, "swizzled operand alignment");

struct Params {
  int M, N, K, act, accumulate, splits, mtiles, ntiles, kblocks, flush;
  float gslope;
  const float* bias;
  float* C;
  int ldc;
  int has_gate;
  unsigned long long* dbg;   // NEMO_B200_GEMM_TIMING: clock stamps of CTA 0
};

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_gemm_tc2(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB,
                                                             const __grid_constant__ CUtensorMap tmC,
                                                             const __grid_constant__ CUtensorMap tmG, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + OFF_BARS;
  auto full = [&](int s) { return bars + 8u * s; };                  // TMA landed                      (tx)
  auto conv = [&](int s) { return bars + 8u * (STAGES + s); };       // A in TMEM, B lo written          (256)
  auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };  // MMAs done with the stage         (commit)
  const uint32_t pfull = bars + 8u * (3 * STAGES);                   // partial accumulator complete     (commit)
  const uint32_t pempty = bars + 8u * (3 * STAGES + 1);              // partial folded into registers    (256)
  auto gbar = [&](int w) { return bars + 8u * (3 * STAGES + 2 + w); };  // gate boxes of epilogue warp w  (tx)
  const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto stamp = [&](int i) {
    if (p.dbg && blockIdx.x == 0) p.dbg[i] = (unsigned long long)clock64();
  };
  if (threadIdx.x == 0) stamp(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(conv(s), 256);
      mbar_init(empty(s), 1);
    }
    mbar_init(pfull, 1);
    mbar_init(pempty, 256);
    for (int w = 0; w < 8; ++w) mbar_init(gbar(w), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32u(tmem_slot);
  if (threadIdx.x == 0) stamp(1);

  const int units = p.mtiles * p.ntiles * p.splits;
  constexpr uint32_t idesc = umma_idesc_tf32(BM, BN, 0, B_MN ? 1 : 0);  // A is K-major in TMEM whatever its source
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
            if (it == 0) stamp(2);
            mbar_arrive_expect_tx(full(s), RAW_BYTES);
            const uint32_t sa = sbase + s * STAGE_BYTES, sb = sa + A_BYTES;
            if (!A_MN) {
              tma_load_2d(sa, &tmA, full(s), kb * BK, mb * BM);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 32; ++i) tma_load_2d(sa + i * (BK * 128), &tmA, full(s), mb * BM + 32 * i, kb * BK);
            }
            if (!B_MN) {
              tma_load_2d(sb, &tmB, full(s), kb * BK, nb * BN);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 32; ++i) tma_load_2d(sb + i * (BK * 128), &tmB, full(s), nb * BN + 32 * i, kb * BK);
            }
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
              if (it == 0) stamp(4);
              if (it == 16) stamp(10);
              if (it == 48) stamp(11);
              tc_fence_after();
              const uint32_t a_hi = tmem_base + TMEM_A + s * (2 * BK), a_lo = a_hi + BK;
              const uint32_t b_hi = sbase + s * STAGE_BYTES + A_BYTES, b_lo = b_hi + B_BYTES;
              // K-major: 64-byte rows, 8-row groups of 512 bytes; MN-major: LBO = stride between 32-wide n blocks
              const uint64_t bh0 = B_MN ? umma_smem_desc(b_hi, BK * 128, 512, LAYOUT_SW128_BASE32B)
                                        : umma_smem_desc(b_hi, 16, 512, LAYOUT_SW64);
              const uint64_t bl0 = B_MN ? umma_smem_desc(b_lo, BK * 128, 512, LAYOUT_SW128_BASE32B)
                                        : umma_smem_desc(b_lo, 16, 512, LAYOUT_SW64);
#pragma unroll
              for (int j = 0; j < BK / UMMA_K; ++j) {
                constexpr uint32_t BSTEP = B_MN ? UMMA_K * 128 : UMMA_K * 4;
                const uint64_t bh = umma_desc_advance(bh0, j * BSTEP), bl = umma_desc_advance(bl0, j * BSTEP);
                umma_tf32_ts(d, a_lo + j * UMMA_K, bh, idesc, (kb > kc || j > 0) ? 1u : 0u);
                umma_tf32_ts(d, a_hi + j * UMMA_K, bl, idesc, 1u);
                umma_tf32_ts(d, a_hi + j * UMMA_K, bh, idesc, 1u);
              }
              umma_commit(empty(s));  // stage (raw tiles, B lo, TMEM A slot) reusable once these MMAs have completed
            }
            umma_commit(pfull);
          }
        }
        stamp(5);
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ A converters (128 threads, thread = row)
    setmaxnreg_dec<56>();
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_A;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int kb0, kb1;
      krange(u % p.splits, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t sa = sbase + s * STAGE_BYTES;
        mbar_wait(full(s), (it / STAGES) & 1);
        if (it == 0 && threadIdx.x == 128) stamp(3);
        uint32_t h[BK], w[BK];
        if (!A_MN) {
          // K-major raw tile: row r = 64 bytes, 16-byte chunk c stored at chunk c ^ ((r >> 1) & 3)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 x = lds128(sa + row * 64 + ((c ^ ((row >> 1) & 3)) << 4));
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float t = trunc_tf32(xs[e]);
              h[4 * c + e] = __float_as_uint(t);
              w[4 * c + e] = __float_as_uint(xs[e] - t);
            }
          }
        } else {
          // MN-major raw tile: [m-block q][k row = 128 bytes of 32 m], 32-byte chunk j stored at chunk j ^ (k & 3)
#pragma unroll
          for (int k = 0; k < BK; ++k) {
            const float x = lds32(sa + q * (BK * 128) + k * 128 + ((((lane >> 3) ^ k) & 3) << 5) + ((lane & 7) << 2));
            const float t = trunc_tf32(x);
            h[k] = __float_as_uint(t);
            w[k] = __float_as_uint(x - t);
          }
        }
        tmem_st16(trow + s * (2 * BK), h);
        tmem_st16(trow + s * (2 * BK) + BK, w);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(conv(s));
      }
    }
  } else if (warp >= 16) {
    // ------------------------------------------------------------------ B converters (128 threads)
    setmaxnreg_dec<40>();
    const int ct = threadIdx.x - 512;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int kb0, kb1;
      krange(u % p.splits, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t hi = sbase + s * STAGE_BYTES + A_BYTES + 16 * ct, lo = hi + B_BYTES;
        mbar_wait(full(s), (it / STAGES) & 1);   // implies the previous round's MMAs have released the stage
#pragma unroll
        for (int bt = 0; bt < 2; ++bt) {
          float4 x[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) x[j] = lds128(hi + 2048 * (4 * bt + j));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 v;
            v.x = x[j].x - trunc_tf32(x[j].x); v.y = x[j].y - trunc_tf32(x[j].y);
            v.z = x[j].z - trunc_tf32(x[j].z); v.w = x[j].w - trunc_tf32(x[j].w);
            sts128(lo + 2048 * (4 * bt + j), v);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(conv(s));
      }
    }
  } else {
    // ------------------------------------------------------------------ accumulate + epilogue (256 threads)
    setmaxnreg_inc<168>();
    const int q = warp & 3, hc = (warp - 8) >> 2;  // TMEM lane quarter, column half
    const uint32_t box0 = sbase + OFF_STAGING + (warp - 8) * (2 * BOX_BYTES);
    const uint32_t rowo = lane * 128;
    uint32_t pc = 0, gphase = 0;
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
        if (threadIdx.x == 256) stamp(6);
      }
      const int m_base = mb * BM + q * 32;
      const int ncol0 = nb * BN + hc * NC;
      if (m_base >= p.M || ncol0 >= p.N) continue;  // warp-uniform
      const bool ab = p.bias && split == 0;
      // this warp's 32 rows x 128 columns leave as four 32 x 32 boxes, two in flight (see gemm_tc.cu)
#pragma unroll
      for (int rd = 0; rd < 2; ++rd) {
        const int n_rd = ncol0 + rd * 64;
        if (n_rd >= p.N) break;  // warp-uniform
        const bool two = n_rd + 32 < p.N;
        if (p.has_gate) {
          if (lane == 0) {
            mbar_arrive_expect_tx(gbar(warp - 8), two ? 2 * BOX_BYTES : BOX_BYTES);
            tma_load_2d(box0, &tmG, gbar(warp - 8), n_rd, m_base);
            if (two) tma_load_2d(box0 + BOX_BYTES, &tmG, gbar(warp - 8), n_rd + 32, m_base);
          }
          mbar_wait(gbar(warp - 8), gphase);
          gphase ^= 1;
        }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int n0 = n_rd + 32 * b;
          if (n0 < p.N) {  // warp-uniform
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const int j0 = rd * 64 + b * 32 + 4 * c4;
              float x[4] = {acc[j0], acc[j0 + 1], acc[j0 + 2], acc[j0 + 3]};
              if (ab) {
#pragma unroll
                for (int k = 0; k < 4; ++k) x[k] += (n0 + 4 * c4 + k < p.N) ? __ldg(p.bias + n0 + 4 * c4 + k) : 0.f;
              }
              if (p.act == 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) x[k] = fmaxf(x[k], 0.f);
              } else if (p.act == 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) x[k] = x[k] > 0.f ? x[k] : 0.01f * x[k];
              }
              const uint32_t a16 = box0 + b * BOX_BYTES + rowo + ((c4 ^ (lane & 7)) << 4);
              if (p.has_gate) {
                const float4 g = lds128(a16);
                x[0] *= g.x > 0.f ? 1.f : p.gslope; x[1] *= g.y > 0.f ? 1.f : p.gslope;
                x[2] *= g.z > 0.f ? 1.f : p.gslope; x[3] *= g.w > 0.f ? 1.f : p.gslope;
              }
              sts128(a16, make_float4(x[0], x[1], x[2], x[3]));
            }
          }
        }
        const bool red = p.splits > 1 || p.accumulate;
        if (red) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tmC, box0, n_rd, m_base);
            if (two) tma_reduce_add_2d(&tmC, box0 + BOX_BYTES, n_rd + 32, m_base);
            bulk_commit();
            bulk_wait_read0();   // the boxes are reused by the next round / tile
          }
          __syncwarp();
        } else {
          // plain store: read the boxes back row-wise -- eight lanes per 128-byte row, four rows per instruction -- and
          // write them with 16-byte stores: every instruction covers four full 128-byte row segments.  (A TMA store
          // drains ~16 bytes per cycle and SM: 8.8k cycles for a 128 x 256 tile, exposed when a CTA owns one tile and
          // competing with the operand loads otherwise; measured, NEMO_B200_GEMM_TIMING.)
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
    if (threadIdx.x == 256) stamp(7);
    if (lane == 0) bulk_wait_all();
    if (threadIdx.x == 256) stamp(8);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(9);
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

// row-major fp32 matrix [outer][inner] with row stride ld
int make_map(CUtensorMap* out, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
             uint32_t box_outer, CUtensorMapSwizzle sw) {
  static std::mutex mu;
  static std::map<std::array<uint64_t, 7>, CUtensorMap> cache;
  const std::array<uint64_t, 7> key = {(uint64_t)(uintptr_t)ptr, inner, outer, ld, box_inner, box_outer, (uint64_t)sw};
  std::lock_guard<std::mutex> lk(mu);
  auto f = cache.find(key);
  if (f != cache.end()) {
    *out = f->second;
    return 0;
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("gemm_tc2: cuTensorMapEncodeTiled entry point unavailable");
    return NEMO_B200_ECUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_tc2: cuTensorMapEncodeTiled failed (%d) for %llu x %llu ld %llu box %u x %u", (int)r,
              (unsigned long long)outer, (unsigned long long)inner, (unsigned long long)ld, box_inner, box_outer);
    return NEMO_B200_ECUDA;
  }
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return 0;
}

template <bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tg, const Params& p,
           int grid, cudaStream_t s) {
  NB_SET_SMEM_ONCE((k_gemm_tc2<A_MN, B_MN>), SMEM_BYTES);
  static const int timing = [] { const char* e = getenv("NEMO_B200_GEMM_TIMING"); return e ? atoi(e) : 0; }();
  if (timing) {   // debug: one launch, synchronise, print CTA 0's timeline (cycles since kernel entry)
    static unsigned long long* dbg = nullptr;
    if (!dbg) NB_CUDA(cudaMalloc(&dbg, 16 * sizeof(unsigned long long)));
    NB_CUDA(cudaMemsetAsync(dbg, 0, 16 * sizeof(unsigned long long), s));
    Params q = p;
    q.dbg = dbg;
    k_gemm_tc2<A_MN, B_MN><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta, tb, tc, tg, q);
    NB_CUDA(cudaStreamSynchronize(s));
    unsigned long long h[16];
    NB_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
    static const char* names[12] = {"entry", "setup done", "first TMA", "first full (A conv)", "first MMA", "last MMA issued",
                                    "last partial folded", "epilogue issued", "stores performed", "all warps done",
                                    "k-block 16 issued", "k-block 48 issued"};
    fprintf(stderr, "[gemm_tc2 timing] M=%d N=%d K=%d grid=%d splits=%d flush=%d:", p.M, p.N, p.K, grid, p.splits, p.flush);
    for (int i = 1; i < 12; ++i) fprintf(stderr, "  %s %lld", names[i], h[i] ? (long long)(h[i] - h[0]) : -1LL);
    fprintf(stderr, "\n");
    return 0;
  }
  k_gemm_tc2<A_MN, B_MN><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta, tb, tc, tg, p);
  NB_LAUNCH_CHECK("gemm_tc2");
}

}  // namespace

namespace nb {

// estimated tensor-pipe cycles of the busiest SM: waves x k-blocks x cycles per k-block (+ a per-tile constant)
double gemm_tc2_cost(int M, int N, int K, int sms, int* splits_out) {
  const int tiles = cdiv(M, BM) * cdiv(N, BN), kblocks = cdiv(K, BK);
  int splits = 1;
  if (tiles < sms) splits = max(1, min(sms / tiles, kblocks / 8));
  if (splits_out) *splits_out = splits;
  const double waves = cdiv(tiles * splits, sms);
  return waves * (cdiv(kblocks, splits) * 828.0 + 4000.0);
}

// C and the gate must be 16-byte addressable (TMA epilogue only); operands as for gemm_tc
bool gemm_tc2_applicable(const float* C, int N, int ldc, const float* gate, int ldg) {
  if (((uintptr_t)C & 15) || (ldc & 3)) return false;
  if ((N & 3) != 0 && ldc - N >= 4) return false;   // 16-byte store granularity: see gemm_tc.cu
  if (gate && (((uintptr_t)gate & 15) || (ldg & 3))) return false;
  return true;
}

int gemm_tc2(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda, int ldb,
             int ldc, int transA, int transB, int act, const float* gate, int ldg, float gslope, int accumulate,
             int force_splits, cudaStream_t s) {
  Params p;
  p.M = M; p.N = N; p.K = K; p.act = act; p.accumulate = accumulate; p.gslope = gslope; p.bias = bias;
  p.has_gate = gate != nullptr;
  p.dbg = nullptr;
  p.C = C; p.ldc = ldc;
  p.mtiles = cdiv(M, BM);
  p.ntiles = cdiv(N, BN);
  p.kblocks = cdiv(K, BK);
  const int tiles = p.mtiles * p.ntiles, sms = num_sms();
  int splits = 1;
  if (act == 0) gemm_tc2_cost(M, N, K, sms, &splits);
  if (force_splits > 0) splits = (act == 0) ? min(force_splits, p.kblocks) : 1;
  p.splits = splits;
  p.flush = 16;   // k-blocks (of 16) accumulated in tensor memory before the fold into fp32 registers: 256 k values
  static const int flush_env = [] { const char* e = getenv("NEMO_B200_GEMM_FLUSH"); return e ? atoi(e) : 0; }();
  if (flush_env > 0) p.flush = 2 * flush_env;
  if (splits > 1 && !accumulate)
    NB_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, s));
  CUtensorMap ta, tb, tc, tg;
  // A: transA=0 -> stored (M,K), K contiguous (K-major, 64-byte rows); transA=1 -> stored (K,M), MN-major
  int rc = transA ? make_map(&ta, A, M, K, lda, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                  : make_map(&ta, A, K, M, lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  // B: transB=1 -> stored (N,K), K-major; transB=0 -> stored (K,N), MN-major
  rc = transB ? make_map(&tb, B, K, N, ldb, BK, BN, CU_TENSOR_MAP_SWIZZLE_64B)
              : make_map(&tb, B, N, K, ldb, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  rc = make_map(&tc, C, N, M, ldc, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  tg = tc;
  if (gate) {
    rc = make_map(&tg, gate, N, M, ldg, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const int grid = min(tiles * splits, sms);
  if (transA) {
    if (transB) return launch<true, false>(ta, tb, tc, tg, p, grid, s);
    return launch<true, true>(ta, tb, tc, tg, p, grid, s);
  }
  if (transB) return launch<false, false>(ta, tb, tc, tg, p, grid, s);
  return launch<false, true>(ta, tb, tc, tg, p, grid, s);
}

}  // namespace nb
