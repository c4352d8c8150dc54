// K7 forward, second generation: 16 frame slots per tensor-core pass.
// Reference arithmetic: human_body_prior/body_model/lbs.py:236-252 (v_posed, T = W.A, v = T.[v_posed;1]) and the
// vertex-to-vertex L1 of vposer_loss (nemo/neural_motion_model.py:2792-2797).
//
// A tcgen05.mma.kind::tf32 costs ~96 issue cycles whatever its N up to ~176 and ~106 at N = 192
// (tools/ubench/mma_pace.cu): the first forward issued 9 MMAs (3xTF32 x 3 k-steps) per 8 frame slots at N = 96 and
// the fused-L1 variant -- two meshes, 4 frames per group -- 18 per 8 frames, more tensor-pipe slots (1 728 cycles)
// than the HBM floor of its two pose-offset streams (1 092).  Here one group is 16 slots (16 frames, or 8 frames x 2
// meshes): T[128 vertices, 16 x 12] = W[128, 24] . Amat[24, 192] is 9 MMAs at N = 192 -- half / a quarter of the MMA
// count per frame.  What makes N = 192 fit twice per SM (2 CTAs, 256 TMEM columns and ~87 KB of shared memory each):
//   * the W tile lives in TENSOR MEMORY (TS-mode A operand, tf32 hi | lo columns), not in shared memory;
//   * the transforms are an MN-major B operand -- k row = joint, 128-byte row = 32 of the (slot, r, c) columns, 32-byte
//     chunk j at j ^ (k & 3) -- which is the memory order of A (N, 24, 12): a writer thread moves float4s (one 16-byte
//     load, two 16-byte stores: raw fp32 as the hi part, x - trunc(x) as the lo part) instead of scalars;
//   * T and Amat are single buffered: the other CTA of the SM fills the tensor pipe meanwhile.
// Third step (H16 = true, the default): the blend on fp16 operand PAIRS, like the adjoint GEMM of the backward
// (skin_bwd2.cu).  The kernel ran into the shared-memory pipe, not HBM: per 16-slot group the transforms cost 36 KB of
// operand stores (hi and lo fp32) and 54 KB of operand fetch (9 MMAs x 6 KB) next to the 24 KB ring.  A kind::f16 MMA
// covers K = 16 in the cycles a tf32 one needs for K = 8, and hi = rn16(x), lo = rn16(x - hi) carries the same 22
// mantissa bits as a tf32 pair: the 24 joints (padded to 32) are 2 k steps = 6 MMAs instead of 9, the operand is 24 KB
// written and 36 KB fetched.  Ranges are known: W in [0, 1] is scaled by 2^8, the transforms by 2^4 (|A| < 4094: SMPL
// transforms are metres), so the lo parts keep an absolute resolution of 2^-29 of a transform entry; the workers undo
// the 2^12 exactly.  NEMO_B200_SKIN_FWD_TF32=1 selects the 3xTF32 blend (measurement switch).
// Roles per CTA (9 warps): 4 workers (thread = vertex = TMEM lane), 3 Amat writers, the MMA issuer, the loader that
// keeps a 2-stage ring of pose offsets (16 slots x 3 components x 512 bytes per stage) in flight -- ONE TMA tensor copy
// per mesh and stage (box 128 vertices x 48 or 24 rows of the (N*3, VP) matrix; rows beyond N land as zeros): a bulk
// copy per row cost the loader warp ~125 cycles each, 6 000 cycles per group, which was the period of the kernel.
#ifndef NB_MBAR_SUSPEND_NS
#define NB_MBAR_SUSPEND_NS 1000000
#endif
#include <cstdlib>
#include <cuda_fp16.h>
#include "common.cuh"
#include "sm100.cuh"
#include "tmap.cuh"

using namespace nb;
using namespace nb::sm100;

namespace nb {
void skin_chunking(int N, int frames_per_group, int ctas_per_sm, double setup, int max_chunks, int& groups_per_chunk,
                   int& nchunks);
}

namespace {

constexpr int TILE_V = 128;
constexpr int NT_V = VP / TILE_V;
constexpr int SL = 16;                       // slots per group
constexpr int NCOL = SL * 12;                // 192 accumulator columns
constexpr int NBLK = NCOL / 32;              // 6 n-blocks of the MN-major operand
constexpr int AM_BLK = NJ * 128;             // one n-block: 24 k rows of 128 bytes
constexpr int AM_PART = NBLK * AM_BLK;       // hi (or lo) part: 18 432 bytes
constexpr int ST_BYTES = SL * 3 * TILE_V * 4;
constexpr int OFF_AM = 0;                    // [hi][lo]
constexpr int OFF_ST = 2 * AM_PART;          // 36 864: two ring stages
constexpr int OFF_RED = OFF_ST + 2 * ST_BYTES;
constexpr int OFF_BARS = OFF_RED + 64;
constexpr int SMEM_BYTES = OFF_BARS + 128 + 1024;
constexpr int TM_T = 0, TM_W = 192, TMEM_COLS = 256;
constexpr int THREADS = 9 * 32;
static_assert(2 * SMEM_BYTES <= 227 * 1024, "two CTAs per SM");
static_assert(OFF_ST % 1024 == 0 && AM_BLK % 1024 == 0, "swizzled operand alignment");

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// ---- fp16-pair blend (H16)
// B operand: MN-major fp16, 128-byte swizzle.  [n block of 64 columns][k group of 8 joints: 1024 B][k row: 128 B = 64 n]
// [16-byte chunk j (8 columns) stored at j ^ (k & 7)]; K = 32 rows (joints 24..31 stay zero)
constexpr int KH = 32;
constexpr int AH_BLK = KH * 128;             // one n block: 4 096 bytes
constexpr int AH_NBLK = NCOL / 64;           // 3
constexpr int AH_PART = AH_NBLK * AH_BLK;    // hi (or lo) part: 12 288 bytes
static_assert(2 * AH_PART <= 2 * AM_PART && AH_BLK % 1024 == 0, "the fp16 operand lives in the tf32 operand's space");
constexpr float W_SCALE = 256.f, A_SCALE = 16.f, T_UNSCALE = 1.f / (256.f * 16.f);
// (a, b) -> fp16 pair: hi = rn16, lo = rn16(x - hi); a in the low half
__device__ __forceinline__ void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem], fp16 operands (two per 32-bit TMEM column: even k in the low half), fp32 accumulator
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

struct Args {
  const float* A;      // (N,24,12)   [MODE 1: mesh a]
  const float* PO;     // (N,3,VP)
  const float* A2;     // MODE 1: mesh b
  const float* PO2;
  const float* base;   // (3,VP)
  const float* Wt;     // (24,VP)
  float* V;            // (N,NV,3)
  uint32_t* signs;     // MODE 1: (N,3,2,VP/32)
  float* l1_sum;
  int N, groups_per_chunk, nchunks;
};

// MODE 0: one mesh, 16 frames per group, vertices written.
// MODE 1: fused L1 of two meshes (slots 0-7 = mesh a, 8-15 = mesh b: 8 frames per group), sign bits written.
template <int MODE, bool H16>
__global__ void __launch_bounds__(THREADS, 2) k_skin_fwd2(const __grid_constant__ CUtensorMap tmPO,
                                                          const __grid_constant__ CUtensorMap tmPO2, const Args a) {
  constexpr int FG = MODE == 0 ? SL : SL / 2;  // frames per group
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + OFF_BARS;
  const uint32_t tfull = bars, tempty = bars + 8;              // T complete (commit) / read by the workers (128)
  const uint32_t amfull = bars + 16, amfree = bars + 24;       // Amat written (96) / read by the MMAs (commit)
  auto sfull = [&](int st) { return bars + 8u * (4 + st); };   // ring stage landed (tx bytes)
  auto sempty = [&](int st) { return bars + 8u * (6 + st); };  // ring stage read by the workers (128)
  const uint32_t tmem_slot = bars + 64;
  const uint32_t sred = sbase + OFF_RED;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int vt = blockIdx.x % NT_V, chunk = blockIdx.x / NT_V;
  const int v0 = vt * TILE_V;
  const int ngroups_total = (a.N + FG - 1) / FG;
  const int g_begin = chunk * a.groups_per_chunk;
  const int ng = min(ngroups_total, g_begin + a.groups_per_chunk) - g_begin;
  const int rot = (vt * 5) % ng;  // vertex tiles walk the chunk's groups in rotated order
  auto group_of = [&](int gi) { const int g = gi + rot; return g_begin + (g >= ng ? g - ng : g); };
  const int last = a.N - 1;
  // frame and mesh of a slot
  auto slot_frame = [&](int n0, int s) { return min(n0 + (MODE == 0 ? s : (s & (FG - 1))), last); };

  if (tid == 0) {
    mbar_init(tfull, 1);
    mbar_init(tempty, 128);
    mbar_init(amfull, 96);
    mbar_init(amfree, 1);
    for (int st = 0; st < 2; ++st) {
      mbar_init(sfull(st), 1);
      mbar_init(sempty(st), 128);
    }
    fence_barrier_init();
  }
  if (warp == 7) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32u(tmem_slot);
  if (tid < TILE_V) {
    const int v = v0 + tid;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + TM_W;
    if (H16) {
      // this vertex's 24 weights x 2^8 as an fp16 pair into its TMEM lane: column c = joints 2c (low half), 2c + 1;
      // 16 hi columns, 16 lo columns (joints 24..31 zero)
      uint32_t h[16], l[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float w0 = 2 * c < NJ ? __ldg(a.Wt + (size_t)(2 * c) * VP + v) : 0.f;
        const float w1 = 2 * c + 1 < NJ ? __ldg(a.Wt + (size_t)(2 * c + 1) * VP + v) : 0.f;
        split_h2(w0 * W_SCALE, w1 * W_SCALE, h[c], l[c]);
      }
      tmem_st16(trow, h);
      tmem_st16(trow + 16, l);
    } else {
      // this vertex's 24 weights, tf32 hi / lo, into its TMEM lane: A operand of the T GEMM (k = joint, 24 of 32 columns)
      uint32_t h[32], l[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float w = k < NJ ? __ldg(a.Wt + (size_t)k * VP + v) : 0.f, hi, lo;
        tf32_split(w, hi, lo);
        h[k] = __float_as_uint(hi);
        l[k] = __float_as_uint(lo);
      }
      tmem_st32(trow, h);
      tmem_st32(trow + 32, l);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t idesc = umma_idesc_tf32(128, NCOL, 0, 1);    // B = Amat, MN-major
  constexpr uint32_t idesc_h = umma_idesc_f16(128, NCOL, 0, 1);   // the same on fp16 pairs

  if (warp >= 4 && warp < 7) {
    // ------------------------------------------------------------------ Amat writers (96 threads)
    // float4 unit u = pt + 96 i of the group's 16 x 72: slot u / 72, joint (u % 72) / 3, quarter u % 3 (coalesced reads
    // of A; the conflict-free store order -- eight lanes on one k row -- was measured here: 2-way store conflicts cost
    // less than the scattered 16-byte reads it needs, 0.140 vs 0.143 ms)
    const int pt = tid - 128;
    if (H16) {
      // unit u = pt + 96 i of the group's 24 chunks (8 accumulator columns each) x 24 joints: joint u % 24, chunk u / 24.
      // The eight lanes of a store phase hold eight consecutive joints (k & 7 all different) of one chunk: eight rows,
      // eight different 16-byte slots of the bank window -- conflict free.  A chunk is two aligned float4s of A
      // (columns 8c .. 8c+3 and 8c+4 .. 8c+7; a slot is 12 columns, so the second may belong to the next slot).
      const int kk = pt % 24, c0 = pt / 24;
      // chunk c = c0 + 4 i: n block c >> 3 = i >> 1, chunk within the block (c & 7) = c0 + 4 (i & 1)
      const uint32_t rowb = sbase + OFF_AM + (kk >> 3) * 1024 + (kk & 7) * 128;
      const uint32_t xo0 = (uint32_t)((c0 ^ kk) & 7) << 4, xo1 = (uint32_t)(((c0 + 4) ^ kk) & 7) << 4;
      auto dst_of = [&](int i) { return rowb + (i >> 1) * AH_BLK + ((i & 1) ? xo1 : xo0); };
      // joints 24..31 of the K = 32 operand: zero, once (the weights there are zero, but 0 x NaN is not)
      for (int i = pt; i < 2 * AH_NBLK * 64; i += 96) {
        const int blk = i / 64, r = i % 64;   // blk: [part][n block], r: 16-byte piece of the 1 KB k group 3
        sts128u(sbase + OFF_AM + (blk / AH_NBLK) * AH_PART + (blk % AH_NBLK) * AH_BLK + 3 * 1024 + r * 16, 0u, 0u, 0u, 0u);
      }
      float4 x[6][2];
      auto fetch = [&](int gi) {
        const int n0 = group_of(gi) * FG;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int n = 8 * c0 + 32 * i + 4 * e, slot = n / 12, off = n - 12 * slot;   // column -> (slot, column of A)
            const float* src = (MODE == 1 && slot >= FG) ? a.A2 : a.A;
            x[i][e] = __ldg(reinterpret_cast<const float4*>(src + (size_t)slot_frame(n0, slot) * 288 + kk * 12 + off));
          }
      };
      fetch(0);
      for (int gi = 0; gi < ng; ++gi) {
        if (gi > 0) mbar_wait(amfree, (gi - 1) & 1);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          uint32_t h[4], l[4];
          split_h2(x[i][0].x * A_SCALE, x[i][0].y * A_SCALE, h[0], l[0]);
          split_h2(x[i][0].z * A_SCALE, x[i][0].w * A_SCALE, h[1], l[1]);
          split_h2(x[i][1].x * A_SCALE, x[i][1].y * A_SCALE, h[2], l[2]);
          split_h2(x[i][1].z * A_SCALE, x[i][1].w * A_SCALE, h[3], l[3]);
          sts128u(dst_of(i), h[0], h[1], h[2], h[3]);
          sts128u(dst_of(i) + AH_PART, l[0], l[1], l[2], l[3]);
        }
        fence_proxy_async_smem();
        mbar_arrive(amfull);
        if (gi + 1 < ng) fetch(gi + 1);
      }
    } else {
    uint32_t dst[12];
    int srcoff[12], sl[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const int u = pt + 96 * i, s = u / 72, j = (u % 72) / 3, q = u % 3;
      const int n = s * 12 + 4 * q, nb = n >> 5, np = n & 31;
      dst[i] = sbase + OFF_AM + nb * AM_BLK + j * 128 + ((((np >> 3) ^ j) & 3) << 5) + ((np >> 2) & 1) * 16;
      srcoff[i] = j * 12 + 4 * q;
      sl[i] = s;
    }
    float4 x[12];
    auto fetch = [&](int gi) {
      const int n0 = group_of(gi) * FG;
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const float* src = (MODE == 1 && sl[i] >= FG) ? a.A2 : a.A;
        x[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)slot_frame(n0, sl[i]) * 288 + srcoff[i]));
      }
    };
    fetch(0);
    for (int gi = 0; gi < ng; ++gi) {
      if (gi > 0) mbar_wait(amfree, (gi - 1) & 1);
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        sts128(dst[i], x[i]);
        sts128(dst[i] + AM_PART, make_float4(x[i].x - trunc_tf32(x[i].x), x[i].y - trunc_tf32(x[i].y),
                                             x[i].z - trunc_tf32(x[i].z), x[i].w - trunc_tf32(x[i].w)));
      }
      fence_proxy_async_smem();
      mbar_arrive(amfull);
      if (gi + 1 < ng) fetch(gi + 1);
    }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ loader
    if (lane == 0) {
      tma_prefetch_desc(&tmPO);
      if (MODE == 1) tma_prefetch_desc(&tmPO2);
      for (int gi = 0; gi < ng; ++gi) {
        const int st = gi & 1;
        if (gi >= 2) mbar_wait(sempty(st), ((gi >> 1) - 1) & 1);
        mbar_arrive_expect_tx(sfull(st), ST_BYTES);
        const int n0 = group_of(gi) * FG;
        const uint32_t dst = sbase + OFF_ST + st * ST_BYTES;
        tma_load_2d(dst, &tmPO, sfull(st), v0, n0 * 3);
        if (MODE == 1) tma_load_2d(dst + ST_BYTES / 2, &tmPO2, sfull(st), v0, n0 * 3);
      }
    }
    __syncwarp();
  } else if (warp == 7) {
    // ------------------------------------------------------------------ T GEMM issuer
    if (lane == 0) {
      const uint32_t d = tmem_base + TM_T;
      const uint32_t wh = tmem_base + TM_W, wl = wh + 32;
      // LBO = stride between n-blocks, SBO = 512 (4 k rows); a k step of 8 joints = 1024 bytes
      const uint64_t mh0 = umma_smem_desc(sbase + OFF_AM, AM_BLK, 512, LAYOUT_SW128_BASE32B);
      const uint64_t ml0 = umma_smem_desc(sbase + OFF_AM + AM_PART, AM_BLK, 512, LAYOUT_SW128_BASE32B);
      // fp16 pairs: LBO = stride between the 64-wide n blocks, SBO = 1024 (8 k rows); a k step of 16 joints = 2048
      // bytes of shared memory and 8 TMEM columns
      const uint64_t gh0 = umma_smem_desc(sbase + OFF_AM, AH_BLK, 1024, LAYOUT_SW128);
      const uint64_t gl0 = umma_smem_desc(sbase + OFF_AM + AH_PART, AH_BLK, 1024, LAYOUT_SW128);
      const uint32_t whh = tmem_base + TM_W, wlh = whh + 16;
      for (int gi = 0; gi < ng; ++gi) {
        mbar_wait(amfull, gi & 1);
        if (gi > 0) mbar_wait(tempty, (gi - 1) & 1);
        tc_fence_after();
        if (H16) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {  // K = 32 (24 joints) = 2 steps of 16
            const uint64_t bh = umma_desc_advance(gh0, ks * 2048), bl = umma_desc_advance(gl0, ks * 2048);
            umma_f16_ts(d, wlh + ks * 8, bh, idesc_h, ks > 0 ? 1u : 0u);
            umma_f16_ts(d, whh + ks * 8, bl, idesc_h, 1u);
            umma_f16_ts(d, whh + ks * 8, bh, idesc_h, 1u);
          }
        } else
#pragma unroll
        for (int j = 0; j < 3; ++j) {  // K = 24 joints = 3 steps of 8
          const uint64_t mh = umma_desc_advance(mh0, j * 1024), ml = umma_desc_advance(ml0, j * 1024);
          umma_tf32_ts(d, wl + j * 8, mh, idesc, j > 0 ? 1u : 0u);
          umma_tf32_ts(d, wh + j * 8, ml, idesc, 1u);
          umma_tf32_ts(d, wh + j * 8, mh, idesc, 1u);
        }
        umma_commit(tfull);
        umma_commit(amfree);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ workers: one thread per vertex
    const int v = v0 + tid;
    const bool real = v < NV;
    float bs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) bs[c] = __ldg(a.base + c * VP + v);
    const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + TM_T;
    const uint32_t st_po = sbase + OFF_ST + tid * 4;
    float l1 = 0.f;
    // one batch = 4 slots = 48 accumulator columns: out[q][r] of slots s0 .. s0 + 3
    auto batch = [&](uint32_t ring, int s0, float (&o)[4][3]) {
      uint32_t t[48];
      tmem_ld16(ta + s0 * 12, t);
      tmem_ld16(ta + s0 * 12 + 16, t + 16);
      tmem_ld16(ta + s0 * 12 + 32, t + 32);
      float po[4][3];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 3; ++c) po[q][c] = lds32(ring + ((s0 + q) * 3 + c) * (TILE_V * 4));
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float p0 = bs[0] + po[q][0], p1 = bs[1] + po[q][1], p2 = bs[2] + po[q][2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const uint32_t* tr = t + q * 12 + 4 * r;
          o[q][r] = fmaf(__uint_as_float(tr[0]), p0,
                         fmaf(__uint_as_float(tr[1]), p1, fmaf(__uint_as_float(tr[2]), p2, __uint_as_float(tr[3]))));
          if (H16) o[q][r] *= T_UNSCALE;   // exact: the accumulator carries 2^8 W . 2^4 A
        }
      }
    };
    for (int gi = 0; gi < ng; ++gi) {
      const int st = gi & 1;
      const int n0 = group_of(gi) * FG;
      const uint32_t ring = st_po + st * ST_BYTES;
      mbar_wait(sfull(st), (gi >> 1) & 1);
      mbar_wait(tfull, gi & 1);
      tc_fence_after();
      if (MODE == 0) {
        // (the three stride-12-byte stores per frame were measured against a transpose through shared memory that writes
        // one full 128-byte line per store: 0.154 ms against 0.140 -- the extra shared-memory traffic costs more)
        // Fast path: the group lies inside the batch and the tile inside the mesh (everything but the last group / the
        // last tile) -- no per-frame or per-thread test between the stores, one base pointer per group and constant
        // frame offsets; with the tests inside, every store re-materialised its descriptor and took two branches.
        const bool inside = (n0 + SL - 1 <= last) && (v0 + TILE_V <= NV);   // block-uniform
        float* const dg = a.V + ((size_t)n0 * NV + v) * 3;
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
          float o[4][3];
          batch(ring, 4 * b, o);
          if (b == 3) {  // the last reads of T and of the ring stage are done
            tc_fence_before();
            mbar_arrive(tempty);
            mbar_arrive(sempty(st));
          }
          float* const db = dg + (size_t)b * (4 * NV * 3);
          if (inside) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              __stcs(db + q * (NV * 3), o[q][0]);
              __stcs(db + q * (NV * 3) + 1, o[q][1]);
              __stcs(db + q * (NV * 3) + 2, o[q][2]);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (real && n0 + 4 * b + q <= last) {
                __stcs(db + q * (NV * 3), o[q][0]);
                __stcs(db + q * (NV * 3) + 1, o[q][1]);
                __stcs(db + q * (NV * 3) + 2, o[q][2]);
              }
            }
          }
        }
      } else {
#pragma unroll 1
        for (int b = 0; b < 2; ++b) {
          float oa[4][3], ob[4][3];
          batch(ring, 4 * b, oa);        // mesh a, frames 4b .. 4b + 3 of the group
          batch(ring, FG + 4 * b, ob);   // mesh b, same frames
          if (b == 1) {
            tc_fence_before();
            mbar_arrive(tempty);
            mbar_arrive(sempty(st));
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int n = n0 + 4 * b + q;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float dlt = (real && n <= last) ? oa[q][c] - ob[q][c] : 0.f;
              l1 += fabsf(dlt);
              const unsigned pm = __ballot_sync(0xffffffffu, dlt > 0.f), nm = __ballot_sync(0xffffffffu, dlt < 0.f);
              if (lane == 0 && n <= last) {
                uint32_t* sg = a.signs + ((size_t)n * 3 + c) * 2 * (VP / 32) + (v >> 5);
                sg[0] = pm;
                sg[VP / 32] = nm;
              }
            }
          }
        }
      }
    }
    if (MODE == 1) {
      l1 = warp_sum(l1);
      if (lane == 0) sts32(sred + 4 * warp, l1);
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (tid == 0) atomicAdd(a.l1_sum, lds32(sred) + lds32(sred + 4) + lds32(sred + 8) + lds32(sred + 12));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 7) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int MODE, bool H16>
int launch(const Args& a, cudaStream_t s) {
  NB_SET_SMEM_ONCE((k_skin_fwd2<MODE, H16>), SMEM_BYTES);
  constexpr int ROWS = MODE == 0 ? SL * 3 : SL * 3 / 2;  // box rows per copy
  CUtensorMap m1, m2;
  if (int rc = make_tmap_2d(&m1, a.PO, 4, VP, (uint64_t)a.N * 3, (uint64_t)VP * 4, TILE_V, ROWS)) return rc;
  m2 = m1;
  if (MODE == 1)
    if (int rc = make_tmap_2d(&m2, a.PO2, 4, VP, (uint64_t)a.N * 3, (uint64_t)VP * 4, TILE_V, ROWS)) return rc;
  k_skin_fwd2<MODE, H16><<<NT_V * a.nchunks, THREADS, SMEM_BYTES, s>>>(m1, m2, a);
  NB_LAUNCH_CHECK("skin_fwd2");
}

bool blend_tf32() {
  static const bool v = [] { const char* e = getenv("NEMO_B200_SKIN_FWD_TF32"); return e && atoi(e) != 0; }();
  return v;
}

}  // namespace

namespace nb {

int skin_fwd2(const float* A, const float* PO, const float* base, const float* Wt, float* V, int N, cudaStream_t s) {
  Args a{};
  a.A = A; a.PO = PO; a.base = base; a.Wt = Wt; a.V = V; a.N = N;
  skin_chunking(N, SL, 2, 1.0, 32, a.groups_per_chunk, a.nchunks);
  return blend_tf32() ? launch<0, false>(a, s) : launch<0, true>(a, s);
}

int skin_l1_fwd2(const float* A_a, const float* PO_a, const float* A_b, const float* PO_b, const float* base,
                 const float* Wt, uint32_t* signs, float* l1_sum, int N, cudaStream_t s) {
  Args a{};
  a.A = A_a; a.PO = PO_a; a.A2 = A_b; a.PO2 = PO_b; a.base = base; a.Wt = Wt; a.signs = signs; a.l1_sum = l1_sum;
  a.N = N;
  skin_chunking(N, SL / 2, 2, 1.0, 32, a.groups_per_chunk, a.nchunks);
  return blend_tf32() ? launch<1, false>(a, s) : launch<1, true>(a, s);
}

}  // namespace nb
