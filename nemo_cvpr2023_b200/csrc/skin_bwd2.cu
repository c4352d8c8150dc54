// K7 backward, second generation: the adjoint GEMM on fp16 operand pairs (tcgen05.mma.kind::f16, K = 16).
// Reference arithmetic: the backward of human_body_prior/body_model/lbs.py:236-252 (T = W.A, v = T.[v_posed;1]):
//   dPO[f,v]    = R[f,v]^T dv[f,v]                         R = rotation part of T[f,v] = sum_j W[v,j] A[f,j]
//   dA[f,j]     = sum_v W[v,j] G[f,v]                      G = dv (x) [v_posed;1]   (3 x 4 per vertex-frame)
//
// What bounded the first generation (skin_tc.cu, per-role counters in profiles/r01_k7_bwd_ablation_v3.txt) was the
// shared-memory pipe: G went through shared memory as tf32 hi / lo (96 KB written by the workers and 96 KB fetched by
// the tensor core per 128 vertices x 8 frames) and the adjoint took 16 MMAs per group, each ~96-106 cycles whatever
// its shape.  A kind::f16 MMA costs the same cycles for K = 16 as a kind::tf32 one for K = 8 (tools/ubench/mma_f16.cu,
// measured: M = 64, N = 192: 106 cycles both), and an fp16 PAIR (hi = rn16(x), lo = rn16(x - hi)) carries the same
// 22 mantissa bits as a tf32 pair.  So:
//   * G is written as fp16 hi | lo: 48 KB per group instead of 96 (half the worker stores, half the operand fetch),
//     8 MMAs per group instead of 16.  Exponent range: every frame pair of a group is scaled by a power of two chosen
//     from max |G| over the tile (the four worker warps of the pair exchange their maxima through shared memory while
//     the dPO stores are in flight), so the largest element sits in [2^13, 2^14) and nothing overflows; elements more
//     than 2^17 below the largest lose relative (not absolute) precision, which is what an fp32 sum does to them
//     anyway.  The drain multiplies the accumulator by the inverse power of two: exact.
//   * W^T (the constant A operand, M = 64: [W_hi^T ; W_lo^T] stacked in the rows, two copies at lane offsets 0 / 16 so
//     two accumulators share their columns) is an fp16 pair too, scaled by 2^8: 64 TMEM columns instead of 128.
//   * the drain writes the W_hi rows and the W_lo rows of the accumulator to two slabs with ONE store instruction and
//     hands both to the bulk-reduce engine: no shuffles (384 per group before).
//   * the shared memory that G no longer needs holds a third input-ring stage.
// The blend T = W.A (needed for dPO only) is unchanged: tf32 hi / lo, W tile in tensor memory, 9 MMAs.
#ifndef NB_MBAR_SUSPEND_NS
#define NB_MBAR_SUSPEND_NS 1000000
#endif
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
constexpr int NT_V = VP / TILE_V;       // 54 vertex tiles
constexpr int FB = 8;                   // frames per group
constexpr int NROW = FB * 12;           // 96 columns of T / of each half of the adjoint accumulator
// The group's transforms as the MN-major tf32 B operand of the T GEMM: k row = joint (128 bytes = 32 of the (f,r,c)
// columns, 32-byte chunk c stored at c ^ (k & 3)), three 32-wide n blocks -- the memory order of A (N,24,12), so a
// writer thread moves float4s: raw fp32 as the hi part (the tensor core ignores the low 13 mantissa bits), x - trunc(x)
// as the lo part.
constexpr int AM_BLK = NJ * 128;        // one n block: 24 k rows
constexpr int AM_PART = 3 * AM_BLK;     // hi (or lo) part: 9 216 bytes

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// TMEM -> registers, shape 16x256b, four repeats along the columns: 16 lanes (starting at the lane of taddr) x 32
// columns.  Thread t receives, for repeat i, v[4i], v[4i+1] = lane t/4, columns 8i + 2(t%4), +1 and v[4i+2], v[4i+3] = lane
// t/4 + 8, same columns.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
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
// power-of-two scale of a frame from the maximum |G| of the tile: E = its biased exponent, clamped so that both the
// scale 2^(140 - E) (largest element -> [2^13, 2^14)) and the drain's factor 2^(E - 148) (undoes it and the 2^8 of W)
// are normal numbers
constexpr int W_SCALE_LOG2 = 8;  // W is handed to the tensor core as fp16 pairs of 2^8 W
__device__ __forceinline__ int scale_exponent(const float4& m) {
  const float mm = fmaxf(fmaxf(m.x, m.y), fmaxf(m.z, m.w));
  const int E = (int)((__float_as_uint(mm) >> 23) & 0xffu);
  return min(max(E, 22), 254);
}

struct Cfg {
  // G: MN-major fp16 B operand, n = [hi (f,r,c) 0..95 | lo 96..191] in three 64-wide blocks, k = vertex:
  // [n block][k group of 8: 1024 B][k row: 128 B = 64 n][16-byte chunk j stored at j ^ (k & 7)]
  static constexpr int GB_BYTES = TILE_V * 128;
  static constexpr int OFF_AM = 0;                       // 2 buffers x [hi][lo] x 9 KB
  static constexpr int OFF_G = OFF_AM + 4 * AM_PART;
  // input ring: THREE TMA tensor copies per stage (a bulk-copy instruction costs its issuing warp ~125 cycles per
  // copy -- the 32 row copies per group of the first generation made the loader warp the slowest role of the kernel):
  //   pose offsets: 2-D tensor (N*3 rows x VP floats), box 128 x 24 = [frame][component][128 vertices]
  //   dV (N, 6890, 3): a frame is 82 680 bytes, NOT a multiple of 16, so the even frames and the odd frames are two
  //   tensors with a row stride of two frames (165 360 bytes); the odd one starts 8 bytes before frame 1 (16-byte
  //   aligned), so an odd frame's data sits two floats into its row.  A box holds at most 256 elements per dimension
  //   and must START on a 16-byte boundary (measured: tools/ubench/tma_box.cu): two boxes per parity (vertices 0..63 |
  //   64..127 of the tile), 192 floats x 4 frames for the even frames, 196 x 4 for the odd ones (both start at the
  //   aligned column, the data is read 8 bytes in; the second box re-reads 4 floats of the first).
  // Rows beyond N and the columns beyond the last vertex read as zero.
  static constexpr int NSTAGE = 3;
  static constexpr int ST_PO_BYTES = FB * 3 * TILE_V * 4;
  static constexpr int DV_EBOX = 4 * 192 * 4, DV_OBOX = 4 * 196 * 4, DV_OBOX_PITCH = 3200;   // 128-byte aligned boxes
  static constexpr int ST_DV_BYTES = 2 * DV_EBOX + 2 * DV_OBOX_PITCH;   // [even: 2 boxes][odd: 2 boxes]
  static constexpr int ST_DV_TX = 2 * DV_EBOX + 2 * DV_OBOX;
  static constexpr int ST_BYTES = ST_PO_BYTES + ST_DV_BYTES;
  static constexpr int OFF_ST = OFF_G + 3 * GB_BYTES;
  // accumulator staging: the group's slab of global dA, [frame][joint][12]; per buffer one slab for the W_hi rows and
  // one for the W_lo rows
  static constexpr int STG_BYTES = FB * 288 * 4;
  static constexpr int OFF_STG = OFF_ST + NSTAGE * ST_BYTES;
  static constexpr int OFF_SC = OFF_STG + 2 * STG_BYTES;  // max |G|: [group & 3][frame][lane quarter]
  static constexpr int OFF_INV = OFF_SC + 512;            // the drain's factor per frame: [group & 3][frame]
  static constexpr int OFF_BARS = OFF_INV + 128;
  static constexpr int SMEM_BYTES = OFF_BARS + 512 + 1024;
  static constexpr int TM_T = 0;        // T: 2 x 96 columns (double buffered: the blend of the next group runs ahead)
  static constexpr int TM_WT = 192;     // [W_hi^T ; W_lo^T] fp16: 64 columns (k = vertex, two per column)
  static constexpr int TM_D = 256;      // dA: 192 columns [. G_hi | . G_lo], TWO M = 64 accumulators (lane offsets 0 / 16)
  static constexpr int TM_W = 448;      // W tile, tf32: 32 hi + 32 lo columns
  static constexpr int TMEM_COLS = 512;
};
static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(AM_BLK % 1024 == 0 && Cfg::OFF_G % 1024 == 0 && Cfg::OFF_ST % 128 == 0 && Cfg::ST_BYTES % 128 == 0 &&
              Cfg::ST_PO_BYTES % 128 == 0 && Cfg::OFF_STG % 16 == 0, "operand / TMA alignment");

struct Args {
  const float* A; const float* PO; const float* base; const float* Wt;
  const float* dV; const uint32_t* signs;
  float* dPO; float* dA;
  float sign_scale;
  // signs only (no dV): |G| = sign_scale * |[p;1]| needs no per-frame maximum -- one power of two takes
  // sign_scale to (32, 64], which keeps every vertex coordinate below ~1000 m inside the fp16 range
  float sc_static, inv_static;
  int N, groups_per_chunk, nchunks;
  unsigned long long* dbg;  // TIMING instantiation only: per-role phase cycle counters
  uint32_t* trace;
  // debug (NEMO_B200_SKIN_ABLATE, with NEMO_B200_SKIN_TIMING=1): 1 no dA reductions, 2 no G stores, 4 no dA MMAs, 8 no dPO stores, 16 no loads,
  // 32 no scale exchange, 64 no accumulator drain, 128 no T MMAs
  int ablate;
};

constexpr int TRACE_CTA = 60, TRACE_NG = 24;
constexpr int FH_ = 2;  // = FH below
constexpr int W_PROD_ = 4 * 8 / FH_, W_ISSUE_ = W_PROD_ + 3, W_DRAIN_ = W_PROD_ + 4, W_LOAD_ = W_PROD_ + 8, W_TISSUE_ = W_PROD_ + 9;
constexpr int FH = FH_;                   // frames of the group per worker warp
constexpr int NWORK = 4 * FB / FH;        // worker warps: warp = 4 * (frame subset of the group) + (TMEM lane quarter)
constexpr int THREADS = (NWORK + 10) * 32;

template <bool HAS_DV, bool HAS_SIGNS, bool TIMING>
__global__ void __launch_bounds__(THREADS, 1) k_skin_bwd2(const __grid_constant__ CUtensorMap tmPO,
                                                          const __grid_constant__ CUtensorMap tmDVe,
                                                          const __grid_constant__ CUtensorMap tmDVo, const Args a) {
  uint32_t tk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = 0;
  int tgi = 0;  // TIMING: the group the role is working on (event trace of one CTA: a.trace[warp][group][event])
  auto tick = [&](int i) {
    if (TIMING) {
      const uint32_t now = (uint32_t)clock();
      tk[i] += now - tprev;
      tprev = now;
      if (blockIdx.x == TRACE_CTA && (threadIdx.x & 31) == 0 && tgi < TRACE_NG)
        a.trace[((threadIdx.x >> 5) * TRACE_NG + tgi) * 8 + i] = now;
    }
  };
  auto tick_flush = [&](int role) {
    if (TIMING && (threadIdx.x & 31) == 0)
      for (int i = 0; i < 8; ++i) atomicAdd(a.dbg + role * 8 + i, (unsigned long long)tk[i]);
  };
  using C_ = Cfg;
  const int abl = TIMING ? a.ablate : 0;  // ablation switches exist in the instrumented instantiation only
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + C_::OFF_BARS;
  auto tfull = [&](int b) { return bars + 8u * b; };            // T[b] complete                     (commit)
  auto tempty = [&](int b) { return bars + 8u * (28 + b); };    // T[b] read by the workers          (all worker threads)
  auto amfree = [&](int b) { return bars + 8u * (2 + b); };     // T GEMM done reading Amat[b]       (commit)
  auto amfull = [&](int b) { return bars + 8u * (26 + b); };    // Amat[b] written                   (3 warps)
  auto d2full = [&](int b) { return bars + 8u * (4 + b); };     // dA[b] complete                    (commit)
  auto d2empty = [&](int b) { return bars + 8u * (6 + b); };    // dA[b] copied to the slabs         (256)
  auto gfull = [&](int kb) { return bars + 8u * (8 + kb); };    // G k block written                 (128)
  auto gfree = [&](int kb) { return bars + 8u * (12 + kb); };   // dA GEMM done reading the k block  (commit)
  auto sfull = [&](int st) { return bars + 8u * (16 + st); };   // input stage landed                (tx bytes)
  auto sempty = [&](int st) { return bars + 8u * (20 + st); };  // input stage read by the workers   (warps)
  const uint32_t tmem_slot = bars + 8u * 24;
  static_assert(C_::NSTAGE <= 4, "barrier slots above");
  // the four quarters of a frame subset have posted their maxima of group gi                          (4 warps)
  auto scposted = [&](int gi, int half) { return bars + 8u * (32 + (gi & 3) * 4 + half); };
  auto scale_slot = [&](int gi, int f) { return sbase + C_::OFF_SC + (uint32_t)(((gi & 3) * FB + f) * 16); };
  auto inv_slot = [&](int gi) { return sbase + C_::OFF_INV + (uint32_t)((gi & 3) * FB * 4); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int vt = blockIdx.x % NT_V, chunk = blockIdx.x / NT_V;
  const int v0 = vt * TILE_V;
  const int ngroups_total = (a.N + FB - 1) / FB;
  const int g_begin = chunk * a.groups_per_chunk;
  const int ng = min(ngroups_total, g_begin + a.groups_per_chunk) - g_begin;
  // vertex tiles walk the chunk's groups in rotated order so their dA reductions do not pile onto the same frames
  const int rot = (vt * 5) % ng;
  auto group_of = [&](int gi) { const int g = gi + rot; return g_begin + (g >= ng ? g - ng : g); };
  const int last = a.N - 1;

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), NWORK);
      mbar_init(amfree(b), 1);
      mbar_init(amfull(b), 3);
      mbar_init(d2full(b), 1);
      mbar_init(d2empty(b), 4);
    }
    for (int kb = 0; kb < 4; ++kb) {
      mbar_init(gfull(kb), NWORK / 4);
      mbar_init(gfree(kb), 1);
    }
    for (int i = 0; i < 16; ++i) mbar_init(bars + 8u * (32 + i), 4);
    for (int st = 0; st < C_::NSTAGE; ++st) {
      mbar_init(sfull(st), 1);
      mbar_init(sempty(st), NWORK);
    }
    fence_barrier_init();
  }
  // warps [W_PROD, W_PROD+3): Amat writers; W_ISSUE: dA GEMM; [W_DRAIN, W_DRAIN+8): accumulator drain, two warps per
  // TMEM lane quarter; W_LOAD: input ring; W_TISSUE: T GEMM
  constexpr int W_PROD = NWORK, W_ISSUE = NWORK + 3, W_DRAIN = NWORK + 4, W_LOAD = NWORK + 8, W_TISSUE = NWORK + 9;
  static_assert(W_DRAIN % 4 == 0, "drain warp W_DRAIN + q reads TMEM lane quarter q");
  if (warp == W_ISSUE) tmem_alloc<C_::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32u(tmem_slot);
  if (tid < TILE_V) {
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    {
      // this vertex's 24 weights, tf32 hi / lo, into its TMEM lane: A operand of the T GEMM
      const int v = v0 + tid;
      uint32_t h[32], l[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float w = k < NJ ? __ldg(a.Wt + (size_t)k * VP + v) : 0.f, hi, lo;
        tf32_split(w, hi, lo);
        h[k] = __float_as_uint(hi);
        l[k] = __float_as_uint(lo);
      }
      tmem_st32(trow + C_::TM_W, h);
      tmem_st32(trow + C_::TM_W + 32, l);
    }
    // A operand of the dA GEMM, M = 64, fp16, column c = vertices 2c (low half) and 2c + 1.  An M = 64 operand /
    // accumulator keeps row r in lane 32 (r / 16) + r % 16 (+ 16 for a second copy; operand and accumulator must use
    // the same lane offset: tools/ubench/mma_m64_ts.cu, mma_f16.cu), so the operand is stored twice and the two
    // accumulators of the double-buffered dA GEMM share their 192 columns.  Rows of TMEM lane quarter q:
    // [W_hi^T joints 6q..6q+5 | 2 zero | W_lo^T joints 6q..6q+5 | 2 zero]: the hi and lo rows of a joint are 8 lanes
    // apart, which is what one thread receives from a 16x256b TMEM load.
    const int wrow_l = lane & 15;                              // row within the quarter (both copies)
    const int wrow_j = 6 * warp + (wrow_l & 7);
    const bool wrow_on = (wrow_l & 7) < 6, wrow_hi = wrow_l < 8;
    const float* wrow = a.Wt + (size_t)min(wrow_j, NJ - 1) * VP + v0;
    constexpr float WS = (float)(1 << W_SCALE_LOG2);
#pragma unroll 1
    for (int c2 = 0; c2 < 2; ++c2) {
      uint32_t x[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + c2 * 64) + i);
        uint32_t h0, l0, h1, l1;
        split_h2(w.x * WS, w.y * WS, h0, l0);
        split_h2(w.z * WS, w.w * WS, h1, l1);
        x[2 * i] = wrow_on ? (wrow_hi ? h0 : l0) : 0u;
        x[2 * i + 1] = wrow_on ? (wrow_hi ? h1 : l1) : 0u;
      }
      tmem_st32(trow + C_::TM_WT + c2 * 32, x);
    }
    tmem_st_wait();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t idesc = umma_idesc_tf32(128, NROW, 0, 1);      // B = Amat, MN-major
  constexpr uint32_t idesc_d = umma_idesc_f16(64, 2 * NROW, 0, 1);  // B = G, MN-major

  if (warp >= W_PROD && warp < W_ISSUE) {
    // ------------------------------------------------------------------ Amat writers (double buffered)
    // float4 unit u = pt + 96 i of the group's 24 joints x 24 quads (quad = frame * 3 + quarter of the 3x4 transform):
    // eight consecutive lanes take eight consecutive quads of ONE joint, i.e. the eight 16-byte slots of one 128-byte
    // k row (in swizzled order): the 16-byte stores of a phase are conflict free
    const int pt = tid - W_PROD * 32;
    uint32_t dst[6];
    int srcoff[6], fr[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int u = pt + 96 * i, j = (u >> 3) % 24, gq = ((u >> 3) / 24) * 8 + (u & 7), f = gq / 3, q = gq % 3;
      const int n = f * 12 + 4 * q, nb = n >> 5, np = n & 31;
      dst[i] = sbase + C_::OFF_AM + nb * AM_BLK + j * 128 + ((((np >> 3) ^ j) & 3) << 5) + ((np >> 2) & 1) * 16;
      srcoff[i] = j * 12 + 4 * q;
      fr[i] = f;
    }
    float4 x[6];
    auto fetch = [&](int gi) {
      const int n0 = group_of(gi) * FB;
#pragma unroll
      for (int i = 0; i < 6; ++i)
        x[i] = __ldg(reinterpret_cast<const float4*>(a.A + (size_t)min(n0 + fr[i], last) * 288 + srcoff[i]));
    };
    fetch(0);
    tprev = (uint32_t)clock();
    for (int gi = 0; gi < ng; ++gi) {
      const int b = gi & 1;
      tgi = gi;
      if (gi > 1) mbar_wait(amfree(b), ((gi >> 1) - 1) & 1);
      tick(0);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const uint32_t d = dst[i] + b * 2 * AM_PART;
        sts128(d, x[i]);
        sts128(d + AM_PART, make_float4(x[i].x - trunc_tf32(x[i].x), x[i].y - trunc_tf32(x[i].y),
                                        x[i].z - trunc_tf32(x[i].z), x[i].w - trunc_tf32(x[i].w)));
      }
      fence_proxy_async_smem();
      mbar_arrive_warp(amfull(b));
      tick(1);
      if (gi + 1 < ng) fetch(gi + 1);
      tick(2);
    }
    if (pt == 0) tick_flush(1);
  } else if (warp == W_TISSUE) {
    // ------------------------------------------------------------------ T GEMM issuer
    // T is double buffered: the blend of group gi + 1 is issued while the workers are still reading T of group gi
    if (lane == 0) {
      const uint32_t wh = tmem_base + C_::TM_W, wl = wh + 32;
      tprev = (uint32_t)clock();
      for (int gi = 0; gi < ng; ++gi) {
        const int b = gi & 1;
        tgi = gi;
        mbar_wait(amfull(b), (gi >> 1) & 1);
        tick(0);
        if (gi > 1) mbar_wait(tempty(b), ((gi >> 1) - 1) & 1);
        tick(1);
        tc_fence_after();
        const uint32_t d = tmem_base + C_::TM_T + b * NROW;
        // LBO = stride between n blocks, SBO = 512 (4 k rows); a k step of 8 joints = 1024 bytes
        const uint64_t mh0 = umma_smem_desc(sbase + C_::OFF_AM + b * 2 * AM_PART, AM_BLK, 512, LAYOUT_SW128_BASE32B);
        if (!(abl & 128))
#pragma unroll
        for (int j = 0; j < 3; ++j) {  // K = 24 joints = 3 steps of 8
          const uint64_t mh = umma_desc_advance(mh0, j * 1024), ml = umma_desc_advance(mh0, AM_PART + j * 1024);
          umma_tf32_ts(d, wl + j * 8, mh, idesc, j > 0 ? 1u : 0u);
          umma_tf32_ts(d, wh + j * 8, ml, idesc, 1u);
          umma_tf32_ts(d, wh + j * 8, mh, idesc, 1u);
        }
        umma_commit(tfull(b));
        umma_commit(amfree(b));
        tick(2);
      }
    }
    __syncwarp();
  } else if (warp == W_ISSUE) {
    // ------------------------------------------------------------------ dA GEMM issuer
    if (lane == 0) {
      // LBO = stride between the 64-wide n blocks, SBO = 1024 (8 k rows); a k step of 16 vertices = 2048 bytes of
      // shared memory and 8 TMEM columns
      const uint64_t gh0 = umma_smem_desc(sbase + C_::OFF_G, C_::GB_BYTES, 1024, LAYOUT_SW128);
      const uint32_t wt = tmem_base + C_::TM_WT;
      tprev = (uint32_t)clock();
      for (int gi = 0; gi < ng; ++gi) {
        const int b = gi & 1;
        const uint32_t use = gi >> 1;
        tgi = gi;
        const uint32_t boff = (uint32_t)(16 * b) << 16;  // accumulator AND operand copy of buffer b
        const uint32_t d = tmem_base + C_::TM_D + boff;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          tick(4 + kb);
          mbar_wait(gfull(kb), gi & 1);
          tick(kb);
          if (kb == 0 && use > 0) mbar_wait(d2empty(b), (use - 1) & 1);
          tc_fence_after();
          if (!(abl & 4))
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_f16_ts(d, wt + boff + kb * 16 + ks * 8, umma_desc_advance(gh0, (kb * 2 + ks) * 2048), idesc_d,
                        (kb | ks) ? 1u : 0u);
          umma_commit(gfree(kb));
        }
        umma_commit(d2full(b));
      }
      tick_flush(2);
    }
    __syncwarp();
  } else if (warp >= W_DRAIN && warp < W_DRAIN + 4) {
    // ------------------------------------------------------------------ accumulator drain: one warp per lane quarter
    // Buffer b of the accumulator lives in lanes [16b, 16b + 16) of every quarter; quarter q carries joints 6q..6q+5:
    // lanes +0..5 = W_hi^T rows (columns [. G_hi | . G_lo]), lanes +8..13 = W_lo^T rows (only . G_hi counts).
    // A 16x256b load hands thread t the lanes t/4 AND t/4 + 8 -- the hi and the lo row of joint 6q + t/4 -- for the
    // columns 8i + 2(t%4), +1: dA[j][f][.] = (hi[. G_hi] + lo[. G_hi]) + hi[. G_lo] is three registers of ONE thread,
    // all 32 threads of the warp carry data of 16 lanes (24 of them useful), and the sums go straight into the
    // group's slab (laid out like global dA (N,24,12)) as 8-byte stores.  ONE thread hands the 9 KB slab to the
    // bulk-reduce engine (cp.reduce.async.bulk .add.f32): the 54 vertex tiles add up in L2.
    const int q = warp - W_DRAIN, dt = q * 32 + lane;
    const int jl = lane >> 2, c2 = 2 * (lane & 3);
    const bool on = jl < 6;
    tprev = (uint32_t)clock();
    for (int gi = 0; gi < ng; ++gi) {
      const int b = gi & 1;
      tgi = gi;
      mbar_wait(d2full(b), (gi >> 1) & 1);
      tick(0);
      tc_fence_after();
      const uint32_t slab = sbase + C_::OFF_STG + b * C_::STG_BYTES;
      if (gi >= 2) {  // the bulk reduction that read this slab two groups ago has left shared memory
        if (dt == 0) bulk_wait_read1();
        asm volatile("bar.sync 2, 128;" ::: "memory");
      }
      tick(1);
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32 + 16 * b) << 16) + C_::TM_D;
      float inv8[FB];
      if (HAS_DV) {
        const float4 iv0 = lds128(inv_slot(gi)), iv1 = lds128(inv_slot(gi) + 16);
        inv8[0] = iv0.x, inv8[1] = iv0.y, inv8[2] = iv0.z, inv8[3] = iv0.w;
        inv8[4] = iv1.x, inv8[5] = iv1.y, inv8[6] = iv1.z, inv8[7] = iv1.w;
      } else {
#pragma unroll
        for (int f = 0; f < FB; ++f) inv8[f] = a.inv_static;
      }
      const uint32_t srow = slab + (6 * q + jl) * 48 + c2 * 4;
      if (!(abl & 64))
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {  // 32 columns of each half per batch
        uint32_t x[16], y[16];
        tmem_ld_16x256b_x4(ta + ch * 32, x);
        tmem_ld_16x256b_x4(ta + NROW + ch * 32, y);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          // columns n, n + 1 with n = 32 ch + 8 i + c2: frame n / 12 is fl or fl + 1
          const int nb = 32 * ch + 8 * i, fl = nb / 12, rem = nb % 12;
          const bool up = rem + c2 >= 12;
          const float inv = up ? inv8[fl + 1 < FB ? fl + 1 : FB - 1] : inv8[fl];
          const float o0 = (__uint_as_float(x[4 * i]) + __uint_as_float(x[4 * i + 2]) + __uint_as_float(y[4 * i])) * inv;
          const float o1 = (__uint_as_float(x[4 * i + 1]) + __uint_as_float(x[4 * i + 3]) + __uint_as_float(y[4 * i + 1])) * inv;
          // slab offset of (frame f, column n - 12 f): 1152 f + 4 (n - 12 f) = 4 n + 1104 f
          if (on) sts64(srow + 4 * nb + 1104 * fl + (up ? 1104 : 0), o0, o1);
        }
      }
      tc_fence_before();
      mbar_arrive_warp(d2empty(b));
      tick(2);
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, 128;" ::: "memory");  // slab complete
      if (dt == 0 && !(abl & 1)) {
        const int n0 = group_of(gi) * FB;
        bulk_reduce_add_f32(a.dA + (size_t)n0 * 288, slab, (uint32_t)min(FB, a.N - n0) * 1152u);
        bulk_commit();
      }
      tick(3);
    }
    if (dt == 0) bulk_wait_all();
    if (warp == W_DRAIN) tick_flush(3);
  } else if (warp == W_LOAD) {
    // ------------------------------------------------------------------ input loader
    if (lane == 0) {
      tma_prefetch_desc(&tmPO);
      if (HAS_DV) {
        tma_prefetch_desc(&tmDVe);
        tma_prefetch_desc(&tmDVo);
      }
      tprev = (uint32_t)clock();
      for (int gi = 0; gi < ng; ++gi) {
        const int st = gi % C_::NSTAGE;
        const uint32_t use = gi / C_::NSTAGE;
        tgi = gi;
        if (use > 0) mbar_wait(sempty(st), (use - 1) & 1);
        tick(0);
        const uint32_t sa = sbase + C_::OFF_ST + st * C_::ST_BYTES;
        const int n0 = group_of(gi) * FB;
        if (abl & 16) {
          mbar_arrive(sfull(st));
          continue;
        }
        mbar_arrive_expect_tx(sfull(st), C_::ST_PO_BYTES + (HAS_DV ? C_::ST_DV_TX : 0));
        tma_load_2d(sa, &tmPO, sfull(st), v0, n0 * 3);
        if (HAS_DV) {
          // (a batch of one frame has no odd row: those copies are sent out of bounds and land as zeros)
          const int ro = a.N > 1 ? n0 / 2 : (1 << 20);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tma_load_2d(sa + C_::ST_PO_BYTES + h * C_::DV_EBOX, &tmDVe, sfull(st), v0 * 3 + h * 192, n0 / 2);
            tma_load_2d(sa + C_::ST_PO_BYTES + 2 * C_::DV_EBOX + h * C_::DV_OBOX_PITCH, &tmDVo, sfull(st), v0 * 3 + h * 192, ro);
          }
        }
        tick(1);
      }
    }
    __syncwarp();
  } else if (warp < NWORK) {
    // ------------------------------------------------------------------ workers: 32 / FH threads per vertex
    // warp = 4 * (frame subset) + (TMEM lane quarter): the warps of a quarter share one warp scheduler and one G k block
    // and take FH frames of the group each.  Fewer, fatter worker warps: the per-group fixed cost of a worker (four
    // barrier waits, the scale exchange, address arithmetic) is paid once per FH frames, and with 18 warps per CTA every
    // thread has 112 registers -- nothing is rematerialised.  The kernel is bound by instruction issue, not by bytes.
    const int wq = warp & 3, half = warp >> 2;
    const int kv = wq * 32 + lane;  // vertex within the tile = k row of G
    const int v = v0 + kv;
    float bs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) bs[c] = __ldg(a.base + c * VP + v);
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    // G addresses.  This thread owns the columns n = 12 FH half .. + 12 FH - 1 of its k row (FH frames x (r, c)) in the
    // hi part and the same + 96 in the lo part: 16-byte chunks of two quads each (quad = dv_r * [p0, p1, p2, 1], quads
    // ordered [f0r0, f0r1, f0r2, f1r0, ...]).  The eight lanes of a store phase hit eight different k rows whose chunk
    // positions are XORed with the row: all eight 16-byte slots of the 128-byte bank window, conflict free.
    constexpr int NCH = 12 * FH / 8;
    uint32_t gh[NCH], gl[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int nh = 12 * FH * half + 8 * i, nl = NROW + nh;
      const uint32_t row = sbase + C_::OFF_G + (kv >> 3) * 1024 + (kv & 7) * 128;
      gh[i] = row + (nh >> 6) * C_::GB_BYTES + ((((nh & 63) >> 3) ^ (kv & 7)) << 4);
      gl[i] = row + (nl >> 6) * C_::GB_BYTES + ((((nl & 63) >> 3) ^ (kv & 7)) << 4);
    }
    const unsigned lbit = 1u << lane;
    const float sp = a.sign_scale, sn = -a.sign_scale;

    uint32_t sA = 0u, sB = 0u;  // this warp's 6 FH sign words, one per lane, prefetched one group ahead
    auto load_signs = [&](int g, uint32_t& w) {
      if (!HAS_SIGNS) return;
      const int n0 = g * FB + half * FH;
      const int fc = lane >> 1;  // word (f*3 + c)*2 + {pos, neg}
      const int n = min(n0 + fc / 3, last);
      w = lane < FH * 6 ? __ldg(a.signs + (((size_t)n * 3 + fc % 3) * 2 + (lane & 1)) * (VP / 32) + (v >> 5)) : 0u;
    };
    const uint32_t st_po = sbase + C_::OFF_ST + (half * FH * 3) * (TILE_V * 4) + kv * 4;
    // frame half * FH + f: row (half * FH + f) / 2 of the even (f even) or odd boxes
    const uint32_t st_dv0 = sbase + C_::OFF_ST + C_::ST_PO_BYTES + (kv >> 6) * C_::DV_EBOX + (half * FH / 2) * 768 + (kv & 63) * 12;
    const uint32_t st_dv1 = sbase + C_::OFF_ST + C_::ST_PO_BYTES + 2 * C_::DV_EBOX + (kv >> 6) * C_::DV_OBOX_PITCH +
                            (half * FH / 2) * 784 + (kv & 63) * 12 + 8;
    // g = this group, gn = the next one in the rotated walk (= g at the end); st / sph = ring stage and its phase
    auto body = [&](int gi, int g, int gn, int st, uint32_t sph, uint32_t sw, uint32_t& swn) {
      const int n0 = g * FB + half * FH;
      const int b = gi & 1;
      tgi = gi;
      load_signs(gn, swn);
      // ---- this group's inputs from the ring
      float p[FH][3], dv[FH][3];
      {
        mbar_wait(sfull(st), sph);
        tick(7);
        const uint32_t so = st * C_::ST_BYTES;
#pragma unroll
        for (int f = 0; f < FH; ++f)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            p[f][c] = lds32(st_po + so + (f * 3 + c) * (TILE_V * 4));
            if (HAS_DV) {  // zero beyond N and beyond vertex 6889
              dv[f][c] = lds32(((f & 1) ? st_dv1 + (f >> 1) * 784 : st_dv0 + (f >> 1) * 768) + so + c * 4);
            } else {
              dv[f][c] = 0.f;
            }
          }
        if (TIMING) {  // force the arrival
          float acc = 0.f;
          for (int f = 0; f < FH; ++f)
            for (int c = 0; c < 3; ++c) acc += p[f][c] + dv[f][c];
          if (acc == 1.2345e-30f) tk[7] += 1;
        }
        mbar_arrive_warp(sempty(st));
      }
      tick(0);
      // ---- upstream gradient of the vertex in each frame
      if (HAS_SIGNS) {
#pragma unroll
        for (int f = 0; f < FH; ++f)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const uint32_t pm = __shfl_sync(0xffffffffu, sw, (f * 3 + c) * 2);
            const uint32_t nm = __shfl_sync(0xffffffffu, sw, (f * 3 + c) * 2 + 1);
            const float s = (pm & lbit) ? sp : ((nm & lbit) ? sn : 0.f);
            dv[f][c] += s;
          }
      }
      // ---- v_posed and this warp's bound on |G| of each frame, handed to the other three quarters
#pragma unroll
      for (int f = 0; f < FH; ++f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) p[f][c] += bs[c];
        if (!HAS_DV) continue;  // static scale
        const float md = fmaxf(fmaxf(fabsf(dv[f][0]), fabsf(dv[f][1])), fabsf(dv[f][2]));
        const float mp = fmaxf(fmaxf(fabsf(p[f][0]), fabsf(p[f][1])), fmaxf(fabsf(p[f][2]), 1.f));
        // non-negative floats order like their bits
        const uint32_t mw = __reduce_max_sync(0xffffffffu, __float_as_uint(md * mp));
        if (lane == 0) sts32(scale_slot(gi, half * FH + f) + 4 * wq, __uint_as_float(mw));
      }
      // posted: an arrival, not a rendezvous -- the other quarters only need this warp to have come THIS far, and they
      // look a whole dPO phase later
      if (HAS_DV) mbar_arrive_warp(scposted(gi, half));
      // ---- dPO = R^T dv from the T rows in TMEM, two frames per batch
      tick(1);
      mbar_wait(tfull(b), (gi >> 1) & 1);
      tick(2);
      tc_fence_after();
      {
        const uint32_t ta = tmem_base + lane_addr + C_::TM_T + b * NROW + half * (FH * 12);
        float* d = a.dPO + (size_t)n0 * (3 * VP) + v;
#pragma unroll
        for (int fb = 0; fb < FH; fb += 2) {
          uint32_t t[24];
          tmem_ld16(ta + fb * 12, t);
          tmem_ld8(ta + fb * 12 + 16, t + 16);
          tmem_ld_wait();
          if (fb + 2 == FH) {
            tc_fence_before();
            mbar_arrive_warp(tempty(b));
          }
#pragma unroll
          for (int f = 0; f < 2; ++f) {
            if (n0 + fb + f <= last && !(abl & 8)) {  // warp-uniform
#pragma unroll
              for (int c = 0; c < 3; ++c)
                __stcs(d + (size_t)(fb + f) * (3 * VP) + c * VP,
                       fmaf(__uint_as_float(t[f * 12 + c]), dv[fb + f][0],
                            fmaf(__uint_as_float(t[f * 12 + 4 + c]), dv[fb + f][1],
                                 __uint_as_float(t[f * 12 + 8 + c]) * dv[fb + f][2])));
            }
          }
        }
      }
      tick(4);
      // ---- the frames' scales: the four quarters have posted their maxima
      float sc[FH];
      if (!HAS_DV) {
#pragma unroll
        for (int f = 0; f < FH; ++f) sc[f] = a.sc_static;
      } else {
        mbar_wait(scposted(gi, half), (gi >> 2) & 1);
        int E[FH];
#pragma unroll
        for (int f = 0; f < FH; ++f) {
          E[f] = scale_exponent(lds128(scale_slot(gi, half * FH + f)));
          sc[f] = __uint_as_float((uint32_t)(127 + 13 + 127 - E[f]) << 23);
        }
        if (wq == 0 && lane == 0) {  // the drain's factors: undo the scale and the 2^8 of W
#pragma unroll
          for (int f = 0; f < FH; f += 2)
            sts64(inv_slot(gi) + 4 * (half * FH + f), __uint_as_float((uint32_t)(E[f] - 13 - W_SCALE_LOG2) << 23),
                  __uint_as_float((uint32_t)(E[f + 1] - 13 - W_SCALE_LOG2) << 23));
        }
      }
      // ---- G for the dA GEMM (the previous group's MMAs over this k block must have finished reading it)
      if (gi > 0) mbar_wait(gfree(wq), (gi - 1) & 1);
      tick(5);
      if (!(abl & 2)) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {  // chunk i = quads 2i, 2i + 1; quad qd = (frame qd / 3, row qd % 3)
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int qd = 2 * i + e, f = qd / 3, r = qd % 3;
            const float ds = dv[f][r] * sc[f];
            split_h2(ds * p[f][0], ds * p[f][1], h[2 * e], l[2 * e]);
            split_h2(ds * p[f][2], ds, h[2 * e + 1], l[2 * e + 1]);
          }
          sts128u(gh[i], h[0], h[1], h[2], h[3]);
          sts128u(gl[i], l[0], l[1], l[2], l[3]);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive_warp(gfull(wq));
      tick(6);
    };
    int g = group_of(0), st = 0;
    uint32_t sph = 0;
    auto next_group = [&](int x) { return x + 1 == g_begin + ng ? g_begin : x + 1; };
    auto advance = [&]() {
      g = next_group(g);
      if (++st == C_::NSTAGE) st = 0, sph ^= 1u;
    };
    load_signs(g, sA);
    tprev = (uint32_t)clock();
    for (int gi = 0; gi < ng; gi += 2) {
      body(gi, g, gi + 1 < ng ? next_group(g) : g, st, sph, sA, sB);
      advance();
      if (gi + 1 < ng) {
        body(gi + 1, g, gi + 2 < ng ? next_group(g) : g, st, sph, sB, sA);
        advance();
      }
    }
    tick_flush(0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_ISSUE) {
    tc_fence_after();
    tmem_dealloc<C_::TMEM_COLS>(tmem_base);
  }
}

struct Maps { CUtensorMap po, dve, dvo; };

template <bool HAS_DV, bool HAS_SIGNS>
int launch(const Maps& m, const Args& a, cudaStream_t s) {
  NB_SET_SMEM_ONCE((k_skin_bwd2<HAS_DV, HAS_SIGNS, false>), Cfg::SMEM_BYTES);
  k_skin_bwd2<HAS_DV, HAS_SIGNS, false><<<NT_V * a.nchunks, THREADS, Cfg::SMEM_BYTES, s>>>(m.po, m.dve, m.dvo, a);
  NB_LAUNCH_CHECK("skin_bwd2");
}

// NEMO_B200_SKIN_TIMING=1 (debug): run the instrumented dV kernel, synchronise and print the phase counters
int launch_timing(const Maps& m, Args a, cudaStream_t s) {
  static unsigned long long* dbg = nullptr;
  static uint32_t* trace = nullptr;
  constexpr int TRACE_WORDS = (THREADS / 32) * TRACE_NG * 8;
  if (!dbg) NB_CUDA(cudaMalloc(&dbg, 40 * sizeof(unsigned long long)));
  if (!trace) NB_CUDA(cudaMalloc(&trace, TRACE_WORDS * 4));
  NB_CUDA(cudaMemsetAsync(dbg, 0, 40 * sizeof(unsigned long long), s));
  NB_CUDA(cudaMemsetAsync(trace, 0, TRACE_WORDS * 4, s));
  a.dbg = dbg;
  a.trace = trace;
  NB_CUDA(cudaFuncSetAttribute(k_skin_bwd2<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Cfg::SMEM_BYTES));
  k_skin_bwd2<true, false, true><<<NT_V * a.nchunks, THREADS, Cfg::SMEM_BYTES, s>>>(m.po, m.dve, m.dvo, a);
  NB_CUDA(cudaStreamSynchronize(s));
  unsigned long long h[40];
  NB_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
  const double groups = (double)NT_V * cdiv(a.N, FB);
  static const char* names[4][8] = {
      {"ring LDS", "signs + max", "wait tfull", "-", "tmem ld + dPO", "scale + wait gfree", "G phase", "wait sfull"},
      {"wait amfree", "Amat store", "fetch", "-", "-", "-", "-", "-"},
      {"wait gfull0", "wait gfull1", "wait gfull2", "wait gfull3", "mma0", "mma1", "mma2", "mma3"},
      {"wait d2full", "wait slab free", "copy + sum", "bulk reduce", "-", "-", "-", "-"}};
  const double per[4] = {(double)NWORK, 1.0, 1.0, 1.0};  // warps per role
  for (int r = 0; r < 4; ++r) {
    fprintf(stderr, "[skin_bwd2 timing] role %d:", r);
    for (int i = 0; i < 8; ++i)
      if (names[r][i][0] != '-') fprintf(stderr, "  %s %.0f", names[r][i], (double)h[r * 8 + i] / groups / per[r]);
    fprintf(stderr, "  (cycles per group per warp)\n");
  }
  // event trace of CTA TRACE_CTA (clock of its SM), groups 8..13: one line per (role warp, group)
  static uint32_t tr[TRACE_WORDS];
  NB_CUDA(cudaMemcpy(tr, trace, sizeof(tr), cudaMemcpyDeviceToHost));
  const int show[] = {0, 3, W_PROD_ - 4, W_PROD_ - 1, W_PROD_, W_ISSUE_, W_DRAIN_, W_DRAIN_ + 3, W_LOAD_, W_TISSUE_};
  const uint32_t t0 = tr[(W_LOAD_ * TRACE_NG + 8) * 8 + 0];
  for (int w : show)
    for (int g = 8; g < 14; ++g) {
      fprintf(stderr, "[skin_bwd2 trace] warp %2d group %2d:", w, g);
      for (int i = 0; i < 8; ++i) {
        const uint32_t v = tr[(w * TRACE_NG + g) * 8 + i];
        if (v) fprintf(stderr, " e%d %6d", i, (int)(v - t0)); else fprintf(stderr, " e%d      -", i);
      }
      fprintf(stderr, "\n");
    }
  return 0;
}

}  // namespace

namespace nb {
int skin_bwd2(const float* A, const float* PO, const float* base, const float* Wt, const float* dV,
              const uint32_t* signs, float sign_scale, float* dPO, float* dA, int N, cudaStream_t s) {
  Args a{};
  a.A = A; a.PO = PO; a.base = base; a.Wt = Wt; a.dV = dV; a.signs = signs;
  a.sign_scale = sign_scale; a.dPO = dPO; a.dA = dA; a.N = N;
  skin_chunking(N, FB, 1, 2.0, 16, a.groups_per_chunk, a.nchunks);
  static const int abl = [] { const char* e = getenv("NEMO_B200_SKIN_ABLATE"); return e ? atoi(e) : 0; }();
  a.ablate = abl;
  {
    int e = 0;
    const float ss = fabsf(sign_scale);
    if (ss > 0.f && isfinite(ss)) frexpf(ss, &e);      // ss = m 2^e, m in [0.5, 1): ss 2^(6 - e) in [32, 64)
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    a.sc_static = ldexpf(1.f, 6 - e);
    a.inv_static = ldexpf(1.f, e - 6 - W_SCALE_LOG2);
  }
  Maps m;
  if (int rc = make_tmap_2d(&m.po, PO, 4, VP, (uint64_t)N * 3, (uint64_t)VP * 4, TILE_V, FB * 3)) return rc;
  m.dve = m.dvo = m.po;
  if (dV) {
    constexpr uint64_t FRAME = (uint64_t)NV * 12;  // 82 680 bytes
    if (int rc = make_tmap_2d(&m.dve, dV, 4, FRAME / 4, (uint64_t)(N + 1) / 2, 2 * FRAME, 192, FB / 2)) return rc;
    if (int rc = N > 1 ? make_tmap_2d(&m.dvo, reinterpret_cast<const char*>(dV) + FRAME - 8, 4, FRAME / 4 + 2,
                                      (uint64_t)N / 2, 2 * FRAME, 196, FB / 2)
                       : make_tmap_2d(&m.dvo, dV, 4, FRAME / 4, 1, 2 * FRAME, 196, FB / 2))
      return rc;
  }
  NB_CUDA(cudaMemsetAsync(dA, 0, sizeof(float) * (size_t)N * 288, s));  // vertex tiles accumulate with reductions
  static const int timing = [] { const char* e = getenv("NEMO_B200_SKIN_TIMING"); return e ? atoi(e) : 0; }();
  if (timing && dV && !signs) return launch_timing(m, a, s);
  if (dV && signs) return launch<true, true>(m, a, s);
  if (dV) return launch<true, false>(m, a, s);
  return launch<false, true>(m, a, s);
}
}  // namespace nb
