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
// Roles per CTA (9 warps): 4 workers (thread = vertex = TMEM lane), 3 Amat writers, the MMA issuer, the loader that
// keeps a 2-stage ring of pose offsets (16 slots x 3 components x 512 bytes per stage) in flight -- ONE TMA tensor copy
// per mesh and stage (box 128 vertices x 48 or 24 rows of the (N*3, VP) matrix; rows beyond N land as zeros): a bulk
// copy per row cost the loader warp ~125 cycles each, 6 000 cycles per group, which was the period of the kernel.
#ifndef NB_MBAR_SUSPEND_NS
#define NB_MBAR_SUSPEND_NS 1000000
#endif
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
template <int MODE>
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
    // this vertex's 24 weights, tf32 hi / lo, into its TMEM lane: A operand of the T GEMM (k = joint, 24 of 32 columns)
    const int v = v0 + tid;
    uint32_t h[32], l[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      float w = k < NJ ? __ldg(a.Wt + (size_t)k * VP + v) : 0.f, hi, lo;
      tf32_split(w, hi, lo);
      h[k] = __float_as_uint(hi);
      l[k] = __float_as_uint(lo);
    }
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + TM_W;
    tmem_st32(trow, h);
    tmem_st32(trow + 32, l);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t idesc = umma_idesc_tf32(128, NCOL, 0, 1);  // B = Amat, MN-major

  if (warp >= 4 && warp < 7) {
    // ------------------------------------------------------------------ Amat writers (96 threads)
    // float4 unit u = pt + 96 i of the group's 16 x 72: slot u / 72, joint (u % 72) / 3, quarter u % 3 (coalesced reads
    // of A; the conflict-free store order -- eight lanes on one k row -- was measured here: 2-way store conflicts cost
    // less than the scattered 16-byte reads it needs, 0.140 vs 0.143 ms)
    const int pt = tid - 128;
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
      for (int gi = 0; gi < ng; ++gi) {
        mbar_wait(amfull, gi & 1);
        if (gi > 0) mbar_wait(tempty, (gi - 1) & 1);
        tc_fence_after();
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
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
          float o[4][3];
          batch(ring, 4 * b, o);
          if (b == 3) {  // the last reads of T and of the ring stage are done
            tc_fence_before();
            mbar_arrive(tempty);
            mbar_arrive(sempty(st));
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int n = n0 + 4 * b + q;
            // (the three stride-12-byte stores were measured against a transpose through shared memory that writes
            // one full 128-byte line per store: 0.154 ms against 0.140 -- the extra shared-memory traffic costs more)
            if (real && n <= last) {  // frame test is warp-uniform
              float* d = a.V + ((size_t)n * NV + v) * 3;
              __stcs(d, o[q][0]);
              __stcs(d + 1, o[q][1]);
              __stcs(d + 2, o[q][2]);
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

template <int MODE>
int launch(const Args& a, cudaStream_t s) {
  NB_SET_SMEM_ONCE(k_skin_fwd2<MODE>, SMEM_BYTES);
  constexpr int ROWS = MODE == 0 ? SL * 3 : SL * 3 / 2;  // box rows per copy
  CUtensorMap m1, m2;
  if (int rc = make_tmap_2d(&m1, a.PO, 4, VP, (uint64_t)a.N * 3, (uint64_t)VP * 4, TILE_V, ROWS)) return rc;
  m2 = m1;
  if (MODE == 1)
    if (int rc = make_tmap_2d(&m2, a.PO2, 4, VP, (uint64_t)a.N * 3, (uint64_t)VP * 4, TILE_V, ROWS)) return rc;
  k_skin_fwd2<MODE><<<NT_V * a.nchunks, THREADS, SMEM_BYTES, s>>>(m1, m2, a);
  NB_LAUNCH_CHECK("skin_fwd2");
}

}  // namespace

namespace nb {

int skin_fwd2(const float* A, const float* PO, const float* base, const float* Wt, float* V, int N, cudaStream_t s) {
  Args a{};
  a.A = A; a.PO = PO; a.base = base; a.Wt = Wt; a.V = V; a.N = N;
  skin_chunking(N, SL, 2, 1.0, 32, a.groups_per_chunk, a.nchunks);
  return launch<0>(a, s);
}

int skin_l1_fwd2(const float* A_a, const float* PO_a, const float* A_b, const float* PO_b, const float* base,
                 const float* Wt, uint32_t* signs, float* l1_sum, int N, cudaStream_t s) {
  Args a{};
  a.A = A_a; a.PO = PO_a; a.A2 = A_b; a.PO2 = PO_b; a.base = base; a.Wt = Wt; a.signs = signs; a.l1_sum = l1_sum;
  a.N = N;
  skin_chunking(N, SL / 2, 2, 1.0, 32, a.groups_per_chunk, a.nchunks);
  return launch<1>(a, s);
}

}  // namespace nb
