// K2 / K5: fp32 GEMM on the 5th-generation tensor cores with the 3xTF32 scheme (fp32-level accuracy),
// replacing the cuBLAS SGEMM calls behind nn.Linear in FCNN / MotionNet (nemo/neural_motion_model.py:58-71,
// 130-148), VPoser (human_body_prior/models/vposer_model.py:63-88), the pose-blend matmul (lbs.py:229-233)
// and their autograd backward (dgrad / wgrad).
//
// Persistent, warp-specialised kernel, one CTA (512 threads) per SM, 128 x 128 x 32 tiles:
//   warp 0       TMA producer    cp.async.bulk.tensor of the raw fp32 A / B boxes into a 4-deep shared-memory ring
//   warps 4-7    A converters    thread = tile row: read the row's 32 k values from the (swizzled) raw tile, split
//                                into tf32 hi / lo and tcgen05.st both into TENSOR MEMORY: the A operand of the
//                                MMAs comes from TMEM, which removes its shared-memory operand traffic (with both
//                                operands in shared memory a 128 x 128 x 8 MMA is fetch-bound at ~50 % of the pipe)
//   warps 2-3    B converters    write lo = x - trunc_tf32(x) into a 2-deep side ring; the raw tile is the hi operand
//                                (the tensor core ignores the low 13 mantissa bits; measured)
//   warp 1       MMA issuer      per 8-wide k step: lo.hi + hi.lo + hi.hi via tcgen05.mma.kind::tf32 (A from TMEM, B
//                                by shared-memory descriptor) into a TMEM partial (2 x 128 columns, double buffered)
//   warps 8-15   accumulate +    tcgen05.ld the partial every `flush` k-blocks and add it into fp32 registers
//                epilogue        (round-to-nearest; the tensor core accumulates with truncation, ~K/8*3*2^-24
//                                relative on long chains), 64 columns per warp; then bias / activation / gate and the
//                                tile leaves through a swizzled shared-memory box and ONE TMA store per 32 x 32 block
//                                (TMA reduce-add for C += and for split-K: the sum happens at the L2).  A per-thread
//                                row store (each lane its own row: 32 cache lines per instruction) cost ~150 cycles
//                                per instruction in the LSU -- 14k cycles per tile against 37k of MMAs at K = 1000
//                                (tools/gemm_probe.py) -- and stalled the MMA warp behind the undrained partial
// A may be K-major (128-byte swizzle) or MN-major (128-byte swizzle with 32-byte atoms; un-transposed by the
// converters), B K-major or MN-major (the only MN-major layout tf32 operands support), so NN / NT / TN / TT all
// run without a transpose pass.  M, N, K tails are handled by TMA zero fill and predicated stores.
#include <stdlib.h>
#include <cuda.h>  // CUtensorMap types; the driver entry point is fetched at run time (no libcuda link)
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

constexpr int BM = 128, BN = 128, BK = 32, UMMA_K = 8;
constexpr int NUM_THREADS = 512;
constexpr int BCONV_THREADS = 64, ACONV_THREADS = 128, ACC_THREADS = 256;
constexpr int NC = BN / 2;                         // accumulator columns per epilogue thread
constexpr int A_BYTES = BM * BK * 4;               // 16 KB
constexpr int B_BYTES = BN * BK * 4;               // 16 KB
constexpr int RAW_STAGES = 4;
constexpr int RAW_BYTES = A_BYTES + B_BYTES;       // [A raw][B raw = hi operand], the bytes TMA delivers per stage
constexpr int STAGE_BYTES = RAW_BYTES + B_BYTES;   // + [B lo]: written by the converters, as deep as the raw ring (a
                                                   // 2-deep lo ring made MMA(k) wait for convert(k) which waited for
                                                   // MMA(k-2): period (convert + MMA) / 2 per k-block, measured)
constexpr int STAGING_BYTES = 8 * 32 * 33 * 4;     // per epilogue warp: one 32 x 32 box (4 KB, 128-byte swizzle) for the TMA
                                                   // store, or a padded 32 x 33 transpose tile on the unaligned path
constexpr int OFF_STAGING = RAW_STAGES * STAGE_BYTES;
constexpr int OFF_BARS = OFF_STAGING + STAGING_BYTES;
constexpr int TMEM_COLS = 512;                     // partials 2 x 128 | A stages 4 x (32 hi + 32 lo)
constexpr int TMEM_A = 2 * BN;
constexpr int SMEM_BYTES = OFF_BARS + 256 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(TMEM_A + RAW_STAGES * 2 * BK <= TMEM_COLS, "tensor memory budget");

struct GemmParams {
  int M, N, K, ldc, act, ldg, accumulate, splits, mtiles, ntiles, kblocks, flush, vec;
  float gslope;
  const float* bias;
  const float* gate;
  float* C;
  unsigned long long* dbg;   // NEMO_B200_GEMM_TIMING: clock stamps of CTA 0 (see gemm_timing_report)
};

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_gemm_tc(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB,
                                                            const __grid_constant__ CUtensorMap tmC,
                                                            const __grid_constant__ CUtensorMap tmG,
                                                            const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + OFF_BARS;
  auto full = [&](int s) { return bars + 8u * s; };                          // TMA landed            (tx)
  auto conv = [&](int s) { return bars + 8u * (RAW_STAGES + s); };           // A in TMEM, B lo ready (192)
  auto empty = [&](int s) { return bars + 8u * (2 * RAW_STAGES + s); };      // MMAs done with stage  (commit)
  auto pfull = [&](int b) { return bars + 8u * (3 * RAW_STAGES + 2 + b); };  // partial complete      (commit)
  auto pempty = [&](int b) { return bars + 8u * (3 * RAW_STAGES + 4 + b); }; // partial drained       (256)
  const uint32_t tmem_slot = bars + 8u * (3 * RAW_STAGES + 6);
  auto gbar = [&](int w) { return bars + 8u * (3 * RAW_STAGES + 8 + w); };   // gate box of epilogue warp w landed (tx)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto stamp = [&](int i) {
    if (p.dbg && blockIdx.x == 0) p.dbg[i] = (unsigned long long)clock64();
  };
  if (threadIdx.x == 0) stamp(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < RAW_STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(conv(s), ACONV_THREADS + BCONV_THREADS);
      mbar_init(empty(s), 1);
    }
    for (int w = 0; w < 8; ++w) mbar_init(gbar(w), 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(pfull(b), 1);
      mbar_init(pempty(b), ACC_THREADS);
    }
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
          const int s = it % RAW_STAGES;
          mbar_wait(empty(s), ((it / RAW_STAGES) & 1) ^ 1);
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
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int kb0, kb1;
        krange(u % p.splits, kb0, kb1);
        for (int kc = kb0; kc < kb1; kc += p.flush, ++pc) {
          const uint32_t pb = pc & 1, pph = (pc >> 1) & 1;
          mbar_wait(pempty(pb), pph ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + pb * BN;
          const int kce = min(kc + p.flush, kb1);
          for (int kb = kc; kb < kce; ++kb, ++it) {
            const int s = it % RAW_STAGES;
            mbar_wait(conv(s), (it / RAW_STAGES) & 1);
            if (it == 0) stamp(4);
            tc_fence_after();
            const uint32_t a_hi = tmem_base + TMEM_A + s * (2 * BK), a_lo = a_hi + BK;
            const uint32_t b_hi = sbase + s * STAGE_BYTES + A_BYTES, b_lo = b_hi + B_BYTES;
            const uint64_t bh0 = B_MN ? umma_smem_desc(b_hi, BK * 128, 512, LAYOUT_SW128_BASE32B)
                                      : umma_smem_desc(b_hi, 16, 1024, LAYOUT_SW128);
            const uint64_t bl0 = B_MN ? umma_smem_desc(b_lo, BK * 128, 512, LAYOUT_SW128_BASE32B)
                                      : umma_smem_desc(b_lo, 16, 1024, LAYOUT_SW128);
#pragma unroll
            for (int j = 0; j < BK / UMMA_K; ++j) {
              constexpr uint32_t BSTEP = B_MN ? UMMA_K * 128 : UMMA_K * 4;
              const uint64_t bh = umma_desc_advance(bh0, j * BSTEP), bl = umma_desc_advance(bl0, j * BSTEP);
              umma_tf32_ts(d, a_lo + j * UMMA_K, bh, idesc, (kb > kc || j > 0) ? 1u : 0u);
              umma_tf32_ts(d, a_hi + j * UMMA_K, bl, idesc, 1u);
              umma_tf32_ts(d, a_hi + j * UMMA_K, bh, idesc, 1u);
            }
            umma_commit(empty(s));    // stage (raw tiles, B lo, TMEM A slot) reusable once these MMAs have completed
          }
          umma_commit(pfull(pb));  // partial accumulator complete
        }
      }
      stamp(5);
    }
    __syncwarp();
  } else if (warp < 4) {
    // ------------------------------------------------------------------ B converters (64 threads)
    const int ct = threadIdx.x - 64;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int kb0, kb1;
      krange(u % p.splits, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % RAW_STAGES;
        const uint32_t hi = sbase + s * STAGE_BYTES + A_BYTES + 16 * ct, lo = hi + B_BYTES;
        constexpr uint32_t CS = 16 * BCONV_THREADS;
        // full(s) of this round implies the previous round's MMAs have released the stage, B lo included
        mbar_wait(full(s), (it / RAW_STAGES) & 1);
        constexpr int NB8 = B_BYTES / 16 / BCONV_THREADS / 8;
#pragma unroll
        for (int bt = 0; bt < NB8; ++bt) {
          float4 x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = lds128(hi + CS * (8 * bt + j));
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 w;
            w.x = x[j].x - trunc_tf32(x[j].x); w.y = x[j].y - trunc_tf32(x[j].y);
            w.z = x[j].z - trunc_tf32(x[j].z); w.w = x[j].w - trunc_tf32(x[j].w);
            sts128(lo + CS * (8 * bt + j), w);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(conv(s));
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ A converters (128 threads, thread = row)
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_A;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int kb0, kb1;
      krange(u % p.splits, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % RAW_STAGES;
        const uint32_t sa = sbase + s * STAGE_BYTES;
        mbar_wait(full(s), (it / RAW_STAGES) & 1);
        if (it == 0 && threadIdx.x == 128) stamp(3);
        uint32_t h[32], w[32];
        if (!A_MN) {
          // K-major raw tile: row r = 128 bytes, 16-byte chunk c stored at chunk (c ^ (r & 7))
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 x = lds128(sa + row * 128 + ((c ^ (row & 7)) << 4));
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float t = trunc_tf32(xs[e]);
              h[4 * c + e] = __float_as_uint(t);
              w[4 * c + e] = __float_as_uint(xs[e] - t);
            }
          }
        } else {
          // MN-major raw tile: [m-block q][k row = 128 bytes of 32 m], 32-byte chunk j stored at chunk (j ^ (k & 3))
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float x = lds32(sa + q * (BK * 128) + k * 128 + ((((lane >> 3) ^ k) & 3) << 5) + ((lane & 7) << 2));
            const float t = trunc_tf32(x);
            h[k] = __float_as_uint(t);
            w[k] = __float_as_uint(x - t);
          }
        }
        // the TMEM slot of this stage was released together with the shared-memory stage (empty -> TMA -> full)
        tmem_st32(trow + s * (2 * BK), h);
        tmem_st32(trow + s * (2 * BK) + BK, w);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(conv(s));
      }
    }
  } else {
    // ------------------------------------------------------------------ accumulate + epilogue (256 threads)
    const int q = warp & 3, hc = (warp - 8) >> 2;  // TMEM lane quarter, column half
    const uint32_t st = sbase + OFF_STAGING + (warp - 8) * (32 * 33 * 4);
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
        const uint32_t pb = pc & 1, pph = (pc >> 1) & 1;
        mbar_wait(pfull(pb), pph);
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + pb * BN + hc * NC;
#pragma unroll
        for (int c = 0; c < NC / 32; ++c) {
          uint32_t v[32];
          tmem_ld16(ta + c * 32, v);
          tmem_ld16(ta + c * 32 + 16, v + 16);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(v[j]);
        }
        tc_fence_before();
        mbar_arrive(pempty(pb));  // partial drained: the MMA warp may overwrite it
        if (threadIdx.x == 256) stamp(6);
      }
      const int m_base = mb * BM + q * 32;
      if (m_base >= p.M) continue;  // warp-uniform
      const int ncol0 = nb * BN + hc * NC;
      if (ncol0 >= p.N) continue;
      if (p.vec) {
        // TMA epilogue: this warp's 32 rows x 64 columns leave as two 32 x 32 boxes.  Box in shared memory: row = 128
        // bytes, 16-byte chunk c of row r at chunk c ^ (r & 7) (SWIZZLE_128B: every quarter warp hits eight distinct
        // chunks).  The gate box (activation backward) arrives the same way and is consumed before it is overwritten.
        const uint32_t box = sbase + OFF_STAGING + (warp - 8) * 4096;
        const uint32_t rowa = box + lane * 128;
        const bool ab = p.bias && split == 0;
#pragma unroll
        for (int c = 0; c < NC / 32; ++c) {
          const int n0 = ncol0 + c * 32;
          if (n0 >= p.N) break;  // warp-uniform
          if (p.gate) {
            if (lane == 0) {
              mbar_arrive_expect_tx(gbar(warp - 8), 4096);
              tma_load_2d(box, &tmG, gbar(warp - 8), n0, m_base);
            }
            mbar_wait(gbar(warp - 8), gphase);
            gphase ^= 1;
          }
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            float x[4] = {acc[c * 32 + 4 * c4], acc[c * 32 + 4 * c4 + 1], acc[c * 32 + 4 * c4 + 2], acc[c * 32 + 4 * c4 + 3]};
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
            const uint32_t a16 = rowa + ((c4 ^ (lane & 7)) << 4);
            if (p.gate) {
              const float4 g = lds128(a16);
              x[0] *= g.x > 0.f ? 1.f : p.gslope; x[1] *= g.y > 0.f ? 1.f : p.gslope;
              x[2] *= g.z > 0.f ? 1.f : p.gslope; x[3] *= g.w > 0.f ? 1.f : p.gslope;
            }
            sts128(a16, make_float4(x[0], x[1], x[2], x[3]));   // own chunk: read (gate) before write, same thread
          }
          if (p.splits > 1 || p.accumulate) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_2d(&tmC, box, n0, m_base);
              bulk_commit();
              bulk_wait_read0();   // the box is reused by the next pass / tile
            }
            __syncwarp();
          } else {
            // plain store: the box is read back row-wise (eight lanes per 128-byte row, four rows per instruction) and
            // written with 16-byte stores -- four full row segments per instruction; see gemm_tc2.cu
            __syncwarp();
            const int nn = n0 + 4 * (lane & 7);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = 4 * i + (lane >> 3);
              const float4 v = lds128(box + r * 128 + (((lane & 7) ^ (r & 7)) << 4));
              if (m_base + r < p.M && nn < p.N)
                *reinterpret_cast<float4*>(p.C + (size_t)(m_base + r) * p.ldc + nn) = v;
            }
            __syncwarp();
          }
        }
        continue;
      }
      // C (or the gate) not 16-byte addressable: transpose through shared memory, coalesced scalar rows
      const int rows = min(32, p.M - m_base);
#pragma unroll
      for (int c = 0; c < NC / 32; ++c) {
        const int n0 = ncol0 + c * 32;
        if (n0 < p.N) {  // warp-uniform
#pragma unroll
          for (int j = 0; j < 32; ++j) sts32(st + 4 * (lane * 33 + j), acc[c * 32 + j]);
          __syncwarp();
          const int n = n0 + lane;
          if (n < p.N) {
            const float bv = (p.bias && split == 0) ? __ldg(p.bias + n) : 0.f;
            for (int i = 0; i < rows; ++i) {
              const size_t m = (size_t)(m_base + i);
              float x = lds32(st + 4 * (i * 33 + lane)) + bv;
              if (p.act == 1) x = fmaxf(x, 0.f);
              else if (p.act == 2) x = x > 0.f ? x : 0.01f * x;
              if (p.gate) x *= (__ldg(p.gate + m * p.ldg + n) > 0.f ? 1.f : p.gslope);
              float* cp = p.C + m * p.ldc + n;
              if (p.splits > 1) atomicAdd(cp, x);
              else if (p.accumulate) *cp += x;
              else *cp = x;
            }
          }
          __syncwarp();
        }
      }
    }
  }
  if (threadIdx.x == 256) stamp(7);
  if (warp >= 8 && lane == 0) bulk_wait_all();   // this warp's TMA stores have been performed
  if (threadIdx.x == 256) stamp(8);
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

// row-major fp32 matrix [outer][inner] with row stride ld; box = 32 x box_outer
int make_map(CUtensorMap* out, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer,
             bool mn_major) {
  static std::mutex mu;
  static std::map<std::array<uint64_t, 6>, CUtensorMap> cache;
  const std::array<uint64_t, 6> key = {(uint64_t)(uintptr_t)ptr, inner, outer, ld, box_outer, (uint64_t)mn_major};
  std::lock_guard<std::mutex> lk(mu);
  auto f = cache.find(key);
  if (f != cache.end()) {
    *out = f->second;
    return 0;
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("gemm_tc: cuTensorMapEncodeTiled entry point unavailable");
    return NEMO_B200_ECUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {32, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) for %llu x %llu ld %llu", (int)r,
              (unsigned long long)outer, (unsigned long long)inner, (unsigned long long)ld);
    return NEMO_B200_ECUDA;
  }
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return 0;
}

template <bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tg,
           const GemmParams& p, int grid, cudaStream_t s) {
  NB_SET_SMEM_ONCE((k_gemm_tc<A_MN, B_MN>), SMEM_BYTES);
  static const int timing = [] { const char* e = getenv("NEMO_B200_GEMM_TIMING"); return e ? atoi(e) : 0; }();
  if (timing) {   // debug: one launch, synchronise, print CTA 0's timeline (cycles since kernel entry)
    static unsigned long long* dbg = nullptr;
    if (!dbg) NB_CUDA(cudaMalloc(&dbg, 16 * sizeof(unsigned long long)));
    NB_CUDA(cudaMemsetAsync(dbg, 0, 16 * sizeof(unsigned long long), s));
    GemmParams q = p;
    q.dbg = dbg;
    k_gemm_tc<A_MN, B_MN><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta, tb, tc, tg, q);
    NB_CUDA(cudaStreamSynchronize(s));
    unsigned long long h[16];
    NB_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
    static const char* names[10] = {"entry", "setup done", "first TMA", "first full (A conv)", "first MMA", "last MMA issued",
                                    "last partial drained", "epilogue issued", "stores performed", "all warps done"};
    fprintf(stderr, "[gemm_tc timing] M=%d N=%d K=%d grid=%d splits=%d flush=%d:", p.M, p.N, p.K, grid, p.splits, p.flush);
    for (int i = 1; i < 10; ++i) fprintf(stderr, "  %s %lld", names[i], h[i] ? (long long)(h[i] - h[0]) : -1LL);
    fprintf(stderr, "\n");
    return 0;
  }
  k_gemm_tc<A_MN, B_MN><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta, tb, tc, tg, p);
  NB_LAUNCH_CHECK("gemm_tc");
}

}  // namespace

namespace nb {

// TMA needs 16-byte aligned bases and row strides; tiny problems stay on the CUDA-core kernel.
bool gemm_tc_applicable(const float* A, const float* B, int M, int N, int K, int lda, int ldb) {
  if (((uintptr_t)A | (uintptr_t)B) & 15) return false;
  if ((lda | ldb) & 3) return false;
  if (K < 8 || M < 16 || N < 8) return false;
  return (double)M * N * K >= 2.0e6;
}

int gemm_tc(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda, int ldb,
            int ldc, int transA, int transB, int act, const float* gate, int ldg, float gslope, int accumulate,
            int force_splits, cudaStream_t s) {
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.act = act; p.ldg = ldg; p.accumulate = accumulate;
  p.gslope = gslope; p.bias = bias; p.gate = gate; p.C = C; p.dbg = nullptr;
  p.mtiles = cdiv(M, BM);
  p.ntiles = cdiv(N, BN);
  p.kblocks = cdiv(K, BK);
  const int tiles = p.mtiles * p.ntiles, sms = num_sms();
  int splits = 1;
  if (act == 0 && tiles < sms) splits = max(1, min(sms / tiles, p.kblocks / 4));
  if (force_splits > 0) splits = (act == 0) ? min(force_splits, p.kblocks) : 1;
  p.splits = splits;
  // k-blocks accumulated in TMEM before the partial is folded into the fp32 registers: short reductions go in
  // one piece (the epilogue then overlaps the next tile), deep ones in chains of 4 x 12 MMAs
  { const int kpu = cdiv(p.kblocks, splits); p.flush = kpu <= 8 ? 8 : 4; }
  static const int flush_env = [] { const char* e = getenv("NEMO_B200_GEMM_FLUSH"); return e ? atoi(e) : 0; }();
  if (flush_env > 0) p.flush = flush_env;   // experiment knob: k-blocks per TMEM partial
  // TMA epilogue: C (and the gate) 16-byte addressable.  TMA stores move whole 16-byte chunks: with N % 4 != 0 the up to
  // three floats [N, N rounded up to 4) of every row receive zeros (+= 0 under reduce-add), so the path is taken only
  // when those floats are the row's trailing padding (ldc - N < 4).
  p.vec = (((uintptr_t)C & 15) == 0 && (ldc & 3) == 0 && ((N & 3) == 0 || ldc - N < 4) &&
           (!gate || (((uintptr_t)gate & 15) == 0 && (ldg & 3) == 0))) ? 1 : 0;
  if (splits > 1 && !accumulate)
    NB_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, s));
  CUtensorMap ta, tb;
  // A: transA=0 -> stored (M,K), K contiguous (K-major); transA=1 -> stored (K,M), M contiguous (MN-major)
  int rc = transA ? make_map(&ta, A, M, K, lda, BK, true) : make_map(&ta, A, K, M, lda, BM, false);
  if (rc) return rc;
  // B: transB=1 -> stored (N,K), K-major; transB=0 -> stored (K,N), N contiguous (MN-major)
  rc = transB ? make_map(&tb, B, K, N, ldb, BN, false) : make_map(&tb, B, N, K, ldb, BK, true);
  if (rc) return rc;
  // epilogue tensor maps: C (store / reduce-add) and the gate, 32 x 32 boxes; unused (copies of A's map) on the
  // unaligned path
  CUtensorMap tc = ta, tg = ta;
  if (p.vec) {
    rc = make_map(&tc, C, N, M, ldc, 32, false);
    if (rc) return rc;
    if (gate) {
      rc = make_map(&tg, gate, N, M, ldg, 32, false);
      if (rc) return rc;
    }
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
