// K7 on the tensor cores: fused linear blend skinning forward (+ fused vertex L1) and backward.
// Reference arithmetic: human_body_prior/body_model/lbs.py:236-252 (v_posed, T = W.A, v = T.[v_posed;1]) and
// nemo/neural_motion_model.py:2792-2797 (vertex-to-vertex L1 of vposer_loss).
//
// With dense skinning weights the blend T[v] = sum_j W[v,j] A[j] costs 288 FMA per vertex-frame and its adjoint
// dA[j] = sum_v W[v,j] (dv (x) [v_posed;1]) another 288: 3x / 6x what the CUDA cores can do in the time HBM needs for
// the frame's streams.  Both contractions therefore run on tcgen05 as 3xTF32 GEMMs inside the memory-bound kernel:
//   T[128 vertices, (frame,12)]   = W_tile[128, 24] . Amat[24, (frame,12)]            accumulator lane = vertex
//   dA^T[(frame,12), joint]       = sum_v G^T[(frame,12), v] . W[v, joint]            G = dv (x) [v_posed;1]
// A CTA owns one 128-vertex tile and walks a chunk of frame groups (8 frames).  The thread that owns a vertex reads
// its 3x4 transforms straight from its TMEM lane, applies them to v_posed = base + pose_offsets (coalesced planar
// reads, prefetched one group ahead) and writes the vertex (forward) or dPO = R^T dv and its 12 G values (backward).
// No per-vertex intermediate reaches HBM.  Joint read-outs are NOT done here: they are linear functionals of the
// mesh and are evaluated without the mesh (readout.cu).
#ifndef NB_MBAR_SLEEP
#define NB_MBAR_SLEEP 32
#endif
#include "common.cuh"
#include "sm100.cuh"

using namespace nb;
using namespace nb::sm100;

namespace {

constexpr int TILE_V = 128;
constexpr int NT_V = VP / TILE_V;       // 54 vertex tiles
constexpr int FB = 8;                   // frames per group (fused L1: 4 frames x 2 meshes)
constexpr int NROW = FB * 12;           // 96 accumulator columns of T / rows of G^T
constexpr int AM_BYTES = NROW * 128;    // one tf32 part of the group's transforms: K-major, row = (slot,r,c), k = joint
constexpr int WT_BYTES = TILE_V * 128;  // one tf32 part of the W tile: K-major, row = vertex, k = joint (24 of 32)
static_assert(VP % TILE_V == 0, "VP must be a multiple of the vertex tile");

// byte offset of element (row, k) in a K-major tile with 128-byte rows (32 fp32 of K) and the 128-byte swizzle
__host__ __device__ __forceinline__ constexpr uint32_t kmajor_sw128(int row, int k) {
  return (uint32_t)(row * 128 + ((((k >> 2) ^ row) & 7) << 4) + ((k & 3) << 2));
}

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// The 96 producer threads turn the 3x4 transforms of a group of 8 frame slots (8 x 24 x 12 floats) into the K-major
// B operand of the T GEMM (row = slot*12 + (r,c), k = joint), tf32 hi / lo.  Thread pt owns elements pt, pt+96,
// pt+192 of every slot's 288, so its shared-memory offsets are loop invariant (two variants: the swizzle phase of
// row = slot*12 + c alternates with the parity of the slot) and its global reads are coalesced.
struct AmatWriter {
  uint32_t off[3][2];
  __device__ __forceinline__ void init(int pt) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int jc = pt + 96 * i, j = jc / 12, c = jc % 12;
      off[i][0] = kmajor_sw128(c, j);
      off[i][1] = kmajor_sw128(12 + c, j) - 12 * 128;
    }
  }
  __device__ __forceinline__ void store(uint32_t am, const float (&x)[FB][3]) const {
#pragma unroll
    for (int s = 0; s < FB; ++s)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float hi, lo;
        tf32_split(x[s][i], hi, lo);
        const uint32_t o = am + s * (12 * 128) + off[i][s & 1];
        sts32(o, hi);
        sts32(o + AM_BYTES, lo);
      }
  }
  // src[s]: the 288 floats of slot s (already clamped to a valid frame)
  template <typename SrcFn>
  __device__ __forceinline__ void write(uint32_t am, int pt, SrcFn src) const {
    float x[FB][3];
#pragma unroll
    for (int s = 0; s < FB; ++s) {
      const float* p = src(s);
#pragma unroll
      for (int i = 0; i < 3; ++i) x[s][i] = __ldg(p + pt + 96 * i);
    }
    store(am, x);
  }
};

// ================================================================================================ forward
// MODE 0: one mesh, 8 frames per group, vertices written.
// MODE 1: fused L1 of two meshes (slots 0-3 = mesh a, 4-7 = mesh b, 4 frames per group), sign bits written.
// 81 KB of shared memory and 256 TMEM columns per CTA: 2 CTAs per SM.
struct FwdCfg {
  static constexpr int OFF_W = 0;                       // [W hi][W lo]
  static constexpr int OFF_AM = 2 * WT_BYTES;           // 2 buffers x [hi][lo]
  // input ring: the pose offsets of NSTAGE groups (8 slots x 3 components x 512 bytes) are in flight as bulk async
  // copies issued by a loader warp: the HBM latency is covered by the ring, not by the workers' registers
  static constexpr int NSTAGE = 2;
  static constexpr int ST_BYTES = FB * 3 * TILE_V * 4;
  static constexpr int OFF_ST = OFF_AM + 4 * AM_BYTES;
  static constexpr int OFF_RED = OFF_ST + NSTAGE * ST_BYTES;
  static constexpr int OFF_BARS = OFF_RED + 256;
  static constexpr int SMEM_BYTES = OFF_BARS + 128 + 1024;
  static constexpr int TMEM_COLS = 256;                 // T: 2 x 128
};
static_assert(2 * FwdCfg::SMEM_BYTES <= 227 * 1024, "two CTAs per SM");
constexpr int FWD_THREADS = 9 * 32;  // 4 workers, 3 Amat writers, T GEMM issuer, loader

struct SkinArgs {
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

template <int MODE>
__global__ void __launch_bounds__(FWD_THREADS, 2) k_skin_tc_fwd(const SkinArgs a) {
  using C_ = FwdCfg;
  constexpr int FG = MODE == 0 ? FB : FB / 2;  // frames per group
  constexpr int NMESH = MODE == 0 ? 1 : 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + C_::OFF_BARS;
  auto tfull = [&](int b) { return bars + 8u * b; };         // T of buffer b complete          (commit)
  auto tempty = [&](int b) { return bars + 8u * (2 + b); };  // workers done reading T[b]       (128)
  auto amfree = [&](int b) { return bars + 8u * (4 + b); };  // MMAs done reading Amat[b]       (commit)
  auto amfull = [&](int b) { return bars + 8u * (7 + b); };  // Amat[b] written                 (96)
  auto sfull = [&](int st) { return bars + 8u * (9 + st); };   // input stage landed            (tx bytes)
  auto sempty = [&](int st) { return bars + 8u * (11 + st); };  // input stage read by the workers (128)
  const uint32_t tmem_slot = bars + 48;
  const uint32_t sred = sbase + C_::OFF_RED;
  static_assert(C_::NSTAGE == 2, "barrier slots above");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int vt = blockIdx.x % NT_V, chunk = blockIdx.x / NT_V;
  const int v0 = vt * TILE_V;
  const int ngroups_total = (a.N + FG - 1) / FG;
  const int g_begin = chunk * a.groups_per_chunk;
  const int ng = min(ngroups_total, g_begin + a.groups_per_chunk) - g_begin;
  // vertex tiles walk the chunk's groups in rotated order (spreads the HBM pages touched at any one time)
  const int rot = (vt * 5) % ng;
  auto group_of = [&](int gi) { const int g = gi + rot; return g_begin + (g >= ng ? g - ng : g); };

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), 128);
      mbar_init(amfree(b), 1);
      mbar_init(amfull(b), 96);
      mbar_init(sfull(b), 1);
      mbar_init(sempty(b), 128);
    }
    fence_barrier_init();
  }
  if (warp == 7) tmem_alloc<C_::TMEM_COLS>(tmem_slot);
  // zero both B-operand buffers once: the k = 24..31 pad columns stay zero for the whole kernel
  for (int i = tid; i < 4 * AM_BYTES / 16; i += FWD_THREADS)
    sts128(sbase + C_::OFF_AM + 16 * i, make_float4(0.f, 0.f, 0.f, 0.f));
  // W tile -> tf32 hi / lo, K-major (row = vertex, k = joint), pad columns zero
  if (tid < TILE_V) {
    const int v = v0 + tid;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      float w = k < NJ ? __ldg(a.Wt + (size_t)k * VP + v) : 0.f, hi, lo;
      tf32_split(w, hi, lo);
      const uint32_t off = sbase + C_::OFF_W + kmajor_sw128(tid, k);
      sts32(off, hi);
      sts32(off + WT_BYTES, lo);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32u(tmem_slot);
  constexpr uint32_t idesc = umma_idesc_tf32(128, NROW, 0, 0);
  const int last = a.N - 1;

  if (warp >= 4) {
    // ------------------------------------------------------------------ producers (96 threads): Amat, double buffered;
    // the transforms of the next group are prefetched into registers while this group's are converted
    if (warp < 7) {
      const int pt = tid - 128;
      AmatWriter aw;
      aw.init(pt);
      float x[FB][3];
      auto fetch = [&](int gi) {
        const int n0 = group_of(gi) * FG;
#pragma unroll
        for (int s = 0; s < FB; ++s) {
          const float* src = (NMESH == 2 && s >= FG) ? a.A2 : a.A;
          const float* p = src + (size_t)min(n0 + (NMESH == 2 ? (s & (FG - 1)) : s), last) * 288 + pt;
#pragma unroll
          for (int i = 0; i < 3; ++i) x[s][i] = __ldg(p + 96 * i);
        }
      };
      fetch(0);
      for (int gi = 0; gi < ng; ++gi) {
        const int b = gi & 1;
        const uint32_t use = gi >> 1;  // how many times buffer b was used before
        if (use > 0) mbar_wait(amfree(b), (use - 1) & 1);
        aw.store(sbase + C_::OFF_AM + b * 2 * AM_BYTES, x);
        fence_proxy_async_smem();
        mbar_arrive(amfull(b));
        if (gi + 1 < ng) fetch(gi + 1);
      }
    } else if (warp == 8) {
      // ---------------------------------------------------------------- input loader: lane = (slot, component)
      for (int gi = 0; gi < ng; ++gi) {
        const int st = gi & 1;
        const uint32_t use = gi >> 1;
        if (use > 0) mbar_wait(sempty(st), (use - 1) & 1);
        if (lane == 0) mbar_arrive_expect_tx(sfull(st), C_::ST_BYTES);
        __syncwarp();
        if (lane < FB * 3) {
          const int s = lane / 3, c = lane % 3;
          const float* P = (NMESH == 2 && s >= FG) ? a.PO2 : a.PO;
          const int n = min(group_of(gi) * FG + (NMESH == 2 ? (s & (FG - 1)) : s), last);
          bulk_g2s(sbase + C_::OFF_ST + st * C_::ST_BYTES + lane * (TILE_V * 4), P + (size_t)n * (3 * VP) + c * VP + v0,
                   TILE_V * 4, sfull(st));
        }
      }
    } else if (lane == 0) {
      // ---------------------------------------------------------------- T GEMM issuer (its own warp: the issue of nine
      // MMAs blocks for most of their execution time)
      for (int gi = 0; gi < ng; ++gi) {
        const int b = gi & 1;
        const uint32_t use = gi >> 1;
        mbar_wait(amfull(b), use & 1);
        if (use > 0) mbar_wait(tempty(b), (use - 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + b * 128;
        const uint64_t wh0 = umma_smem_desc(sbase + C_::OFF_W, 16, 1024, LAYOUT_SW128);
        const uint64_t mh0 = umma_smem_desc(sbase + C_::OFF_AM + b * 2 * AM_BYTES, 16, 1024, LAYOUT_SW128);
#pragma unroll
        for (int j = 0; j < 3; ++j) {  // K = 24 = 3 steps of 8
          const uint64_t wh = umma_desc_advance(wh0, j * 32), wl = umma_desc_advance(wh0, WT_BYTES + j * 32);
          const uint64_t mh = umma_desc_advance(mh0, j * 32), ml = umma_desc_advance(mh0, AM_BYTES + j * 32);
          umma_tf32(d, wl, mh, idesc, j > 0 ? 1u : 0u);
          umma_tf32(d, wh, ml, idesc, 1u);
          umma_tf32(d, wh, mh, idesc, 1u);
        }
        umma_commit(tfull(b));
        umma_commit(amfree(b));
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
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float l1 = 0.f;
    const uint32_t st_po = sbase + C_::OFF_ST + tid * 4;
    auto body = [&](int gi) {
      const int b = gi & 1;
      const int n0 = group_of(gi) * FG;
      // this group's pose offsets from the ring
      float po[FB][3];
      mbar_wait(sfull(b), (gi >> 1) & 1);
#pragma unroll
      for (int s = 0; s < FB; ++s)
#pragma unroll
        for (int c = 0; c < 3; ++c) po[s][c] = lds32(st_po + b * C_::ST_BYTES + (s * 3 + c) * (TILE_V * 4));
      mbar_arrive(sempty(b));
      mbar_wait(tfull(b), (gi >> 1) & 1);
      tc_fence_after();
      const uint32_t ta = tmem_base + lane_addr + b * 128;
      float oa[FB / 2][3];  // MODE 1: mesh a of the group, kept until mesh b arrives in the second batch
#pragma unroll
      for (int h = 0; h < 2; ++h) {  // 4 slots = 48 accumulator columns per batch of loads
        uint32_t t[48];
        tmem_ld16(ta + h * 48, t);
        tmem_ld16(ta + h * 48 + 16, t + 16);
        tmem_ld16(ta + h * 48 + 32, t + 32);
        tmem_ld_wait();
        if (h == 1) {
          tc_fence_before();
          mbar_arrive(tempty(b));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int s = h * 4 + q;
          const float p0 = bs[0] + po[s][0], p1 = bs[1] + po[s][1], p2 = bs[2] + po[s][2];
          float o[3];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const uint32_t* tr = t + q * 12 + 4 * r;
            o[r] = fmaf(__uint_as_float(tr[0]), p0,
                        fmaf(__uint_as_float(tr[1]), p1, fmaf(__uint_as_float(tr[2]), p2, __uint_as_float(tr[3]))));
          }
          if (MODE == 0) {
            if (real && n0 + s <= last) {  // frame test is warp-uniform
              float* d = a.V + ((size_t)(n0 + s) * NV + v) * 3;
              __stcs(d, o[0]);
              __stcs(d + 1, o[1]);
              __stcs(d + 2, o[2]);
            }
          } else if (h == 0) {
            oa[q][0] = o[0]; oa[q][1] = o[1]; oa[q][2] = o[2];
          } else {
            const int n = n0 + q;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float dlt = (real && n <= last) ? oa[q][c] - o[c] : 0.f;
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
    };
    for (int gi = 0; gi < ng; ++gi) body(gi);
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
    tmem_dealloc<C_::TMEM_COLS>(tmem_base);
  }
}

// ================================================================================================ backward
// Per vertex tile and group of 8 frames, two GEMMs on tcgen05, BOTH with their constant operand in TENSOR MEMORY (so
// the only shared-memory operand traffic is the per-frame data):
//   T[v, (f,12)]    = W[v, :] . Amat[:, (f,12)]        A operand = W tile (lane = vertex, tf32 hi / lo columns)
//   dA[j, (f,12)]   = sum_v W^T[j, v] . G[v, (f,12)]   A operand = [W_hi^T ; W_lo^T] stacked in the lanes (hi: lanes
//                                                      0..23, lo: lanes 32..55), K = the 128 vertices of the tile
// G[v,(f,r,c)] = dv[f,v,r] * [v_posed;1][c]: the worker thread of a vertex scatters its 12 values per frame into the
// K-major B operand (row = (f,r,c), k = vertex: a warp writes one full swizzled 128-byte row per store, conflict
// free), raw fp32 as the hi part (the tensor core ignores the low 13 mantissa bits) and G - trunc(G) as the lo part.
// (An MN-major G -- 16-byte stores per thread -- was measured: tf32 MN-major operands are fetched at ~32 bytes per
// clock, which made the dA GEMM the bottleneck.)  Stacking W hi / lo in the accumulator lanes makes 3xTF32 two MMAs
// per k step and G is fetched from shared memory exactly once per part:
//   D[0:24]  += W_hi^T.G_hi + W_hi^T.G_lo        D[32:56] += W_lo^T.G_hi (+ W_lo^T.G_lo)       dA = D[j] + D[32+j]
// G is synchronised per 32-vertex k block (the four worker warps of a TMEM lane quarter share one k block), so the
// MMAs of one block overlap the G writes of the others.
// Roles (23 warps): 16 workers; 3 Amat producers (+ T GEMM issue); 1 dA GEMM issuer; 2 drain warps (TMEM lane
// quarters 0 / 1: the W_lo half goes through a shared staging tile, the W_hi warp adds it and reduces into global
// dA (N,24,12) with 16-byte vector reductions -- the 54 vertex tiles add up); 1 loader feeding the input ring.
struct BwdCfg {
  // G: MN-major B operand, n = [hi (f,r,c) 0..95 | lo 96..191] in six 32-wide blocks, k = vertex:
  // [n block q][k row = 128 bytes of 32 n], 32-byte chunk j of a row stored at chunk (j ^ (k & 3))
  static constexpr int GQ_BYTES = TILE_V * 128;          // one n block: 128 k rows
  static constexpr int GT_LO = 3 * GQ_BYTES;             // n blocks 3..5: the lo part
  static constexpr int OFF_AM = 0;                       // 2 buffers x [hi][lo] x 12 KB
  static constexpr int OFF_G = OFF_AM + 4 * AM_BYTES;    // 4 k blocks x 24 KB
  // input ring: the pose offsets (planar: 3 x 512 bytes per frame) and the upstream vertex gradient (interleaved xyz:
  // 1536 bytes per frame) of NSTAGE frame groups are in flight as 32 bulk async copies per group, one per lane of the
  // loader warp -- the HBM latency is covered by the ring, not by registers.  The frame stride of dV (82 680 bytes) is
  // only 8-byte aligned: an odd frame is copied from 8 bytes earlier (+16 bytes) and read at an 8-byte offset.
  static constexpr int NSTAGE = 2;
  static constexpr int ST_PO_BYTES = FB * 3 * TILE_V * 4;
  static constexpr int DV_SLOT = TILE_V * 12 + 16;
  static constexpr int ST_DV_BYTES = FB * DV_SLOT;
  static constexpr int ST_BYTES = ST_PO_BYTES + ST_DV_BYTES;
  static constexpr int OFF_ST = OFF_G + 6 * GQ_BYTES;
  // accumulator staging: the group's slab of global dA, [frame][joint][12], double buffered
  static constexpr int STG_BYTES = FB * 288 * 4;
  static constexpr int OFF_STG = OFF_ST + NSTAGE * ST_BYTES;
  static constexpr int OFF_BARS = OFF_STG + 2 * STG_BYTES;
  static constexpr int SMEM_BYTES = OFF_BARS + 256 + 1024;
  static constexpr int TM_T = 0;        // T: 96 columns
  static constexpr int TM_WT = 96;      // [W_hi^T ; W_lo^T]: 128 columns (k = vertex)
  static constexpr int TM_D = 224;      // dA: 192 columns [. G_hi | . G_lo], TWO M = 64 accumulators (lane offsets 0 / 16)
  static constexpr int TM_W = 416;      // W tile: 32 hi + 32 lo columns
  static constexpr int TMEM_COLS = 512;
};
static_assert(BwdCfg::SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(BwdCfg::OFF_G % 1024 == 0, "swizzled operand alignment");

struct SkinBwdArgs {
  const float* A; const float* PO; const float* base; const float* Wt;
  const float* dV; const uint32_t* signs;
  float* dPO; float* dA;
  float sign_scale;
  int N, groups_per_chunk, nchunks;
  unsigned long long* dbg;  // TIMING instantiation only: per-role phase cycle counters
  int ablate;  // debug (NEMO_B200_SKIN_ABLATE): 1 no dA reductions, 2 no G stores, 4 no dA MMAs, 8 no dPO stores, 16 no loads
};

constexpr int NWORK = 16;                 // worker warps: warp = 4 * (frame pair of the group) + (TMEM lane quarter)
constexpr int BWD_THREADS = (NWORK + 14) * 32;

template <bool HAS_DV, bool HAS_SIGNS, bool TIMING = false>
__global__ void __launch_bounds__(BWD_THREADS, 1) k_skin_tc_bwd(const SkinBwdArgs a) {
  uint32_t tk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = 0;
  auto tick = [&](int i) {
    if (TIMING) {
      const uint32_t now = (uint32_t)clock();
      tk[i] += now - tprev;
      tprev = now;
    }
  };
  auto tick_flush = [&](int role) {
    if (TIMING && (threadIdx.x & 31) == 0)
      for (int i = 0; i < 8; ++i) atomicAdd(a.dbg + role * 8 + i, (unsigned long long)tk[i]);
  };
  using C_ = BwdCfg;
  constexpr int THREADS = BWD_THREADS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + C_::OFF_BARS;
  const uint32_t tfull = bars, tempty = bars + 8;               // T complete (commit) / read by the workers (512)
  auto amfree = [&](int b) { return bars + 8u * (2 + b); };     // T GEMM done reading Amat[b]       (commit)
  auto amfull = [&](int b) { return bars + 8u * (26 + b); };    // Amat[b] written                   (96)
  auto d2full = [&](int b) { return bars + 8u * (4 + b); };     // dA[b] complete                    (commit)
  auto d2empty = [&](int b) { return bars + 8u * (6 + b); };    // dA[b] copied to the staging tiles (128)
  auto gfull = [&](int kb) { return bars + 8u * (8 + kb); };    // G k block written                 (128)
  auto gfree = [&](int kb) { return bars + 8u * (12 + kb); };   // dA GEMM done reading the k block  (commit)
  auto sfull = [&](int st) { return bars + 8u * (16 + st); };   // input stage landed        (tx bytes [+ 32 cp.async])
  auto sempty = [&](int st) { return bars + 8u * (20 + st); };  // input stage read by the workers   (512)
  const uint32_t tmem_slot = bars + 8u * 24;
  static_assert(C_::NSTAGE <= 4, "barrier slots above");

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
    mbar_init(tfull, 1);
    mbar_init(tempty, NWORK * 32);
    for (int b = 0; b < 2; ++b) {
      mbar_init(amfree(b), 1);
      mbar_init(amfull(b), 96);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(d2full(b), 1);
      mbar_init(d2empty(b), 256);
    }
    for (int kb = 0; kb < 4; ++kb) {
      mbar_init(gfull(kb), NWORK * 8);
      mbar_init(gfree(kb), 1);
    }
    for (int st = 0; st < C_::NSTAGE; ++st) {
      mbar_init(sfull(st), 1);
      mbar_init(sempty(st), NWORK * 32);
    }
    fence_barrier_init();
  }
  // warps [W_PROD, W_PROD+3): Amat writers; W_ISSUE: dA GEMM; [W_DRAIN, W_DRAIN+4): accumulator drain, one warp per
  // TMEM lane quarter; W_LOAD: input ring; W_TISSUE: T GEMM
  constexpr int W_PROD = NWORK, W_ISSUE = NWORK + 3, W_DRAIN = NWORK + 4, W_LOAD = NWORK + 12, W_TISSUE = NWORK + 13;
  static_assert(W_DRAIN % 4 == 0, "drain warp W_DRAIN + q reads TMEM lane quarter q");
  if (warp == W_ISSUE) tmem_alloc<C_::TMEM_COLS>(tmem_slot);
  // zero Amat: its pad columns k = 24..31 stay zero for the whole kernel (G is fully rewritten every group)
  for (int i = tid; i < 4 * AM_BYTES / 16; i += THREADS)
    sts128(sbase + C_::OFF_AM + 16 * i, make_float4(0.f, 0.f, 0.f, 0.f));
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
    // A operand of the dA GEMM, M = 64, column = vertex.  An M = 64 operand / accumulator keeps row r in lane
    // 32 (r / 16) + r % 16 (+ 16 for a second copy: measured, tools/ubench/mma_m64*.cu; operand and accumulator must
    // use the same lane offset), so the operand is stored twice and the two accumulators of the double-buffered dA
    // GEMM share their 192 columns.  Rows of TMEM lane quarter q: [W_hi^T joints 6q..6q+5 | W_lo^T joints 6q..6q+5 |
    // 4 zero] -- the hi and lo parts of a joint sit six lanes apart in ONE quarter, so the drain warp of the quarter
    // adds them with a shuffle.
    const int wrow_l = lane & 15;                              // row within the quarter (both copies)
    const int wrow_j = 6 * warp + (wrow_l < 6 ? wrow_l : wrow_l - 6);
    const bool wrow_on = wrow_l < 12, wrow_hi = wrow_l < 6;
    const float* wrow = a.Wt + (size_t)min(wrow_j, NJ - 1) * VP + v0;
#pragma unroll 1
    for (int c4 = 0; c4 < 4; ++c4) {
      uint32_t x[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + c4 * 32) + i);
        const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float hi, lo;
          tf32_split(ws[e], hi, lo);
          x[4 * i + e] = wrow_on ? __float_as_uint(wrow_hi ? hi : lo) : 0u;
        }
      }
      tmem_st32(trow + C_::TM_WT + c4 * 32, x);
    }
    tmem_st_wait();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t idesc = umma_idesc_tf32(128, NROW, 0, 0);
  constexpr uint32_t idesc_d = umma_idesc_tf32(64, 2 * NROW, 0, 1);  // B = G, MN-major

  if (warp >= W_PROD && warp < W_ISSUE) {
    // ------------------------------------------------------------------ Amat writers (double buffered)
    const int pt = tid - W_PROD * 32;
    AmatWriter aw;
    aw.init(pt);
    float x[FB][3];
    auto fetch = [&](int gi) {
      const int n0 = group_of(gi) * FB;
#pragma unroll
      for (int s = 0; s < FB; ++s) {
        const float* p = a.A + (size_t)min(n0 + s, last) * 288 + pt;
#pragma unroll
        for (int i = 0; i < 3; ++i) x[s][i] = __ldg(p + 96 * i);
      }
    };
    fetch(0);
    tprev = (uint32_t)clock();
    for (int gi = 0; gi < ng; ++gi) {
      const int b = gi & 1;
      if (gi > 1) mbar_wait(amfree(b), ((gi >> 1) - 1) & 1);
      tick(0);
      aw.store(sbase + C_::OFF_AM + b * 2 * AM_BYTES, x);
      fence_proxy_async_smem();
      mbar_arrive(amfull(b));
      tick(1);
      if (gi + 1 < ng) fetch(gi + 1);
      tick(2);
    }
    if (pt == 0) tick_flush(1);
  } else if (warp == W_TISSUE) {
    // ------------------------------------------------------------------ T GEMM issuer
    // T of group gi + 1 is issued as soon as every worker has read T of group gi, i.e. a whole group ahead of its use
    if (lane == 0) {
      const uint32_t d = tmem_base + C_::TM_T;
      const uint32_t wh = tmem_base + C_::TM_W, wl = wh + 32;
      for (int gi = 0; gi < ng; ++gi) {
        const int b = gi & 1;
        mbar_wait(amfull(b), (gi >> 1) & 1);
        if (gi > 0) mbar_wait(tempty, (gi - 1) & 1);
        tc_fence_after();
        const uint64_t mh0 = umma_smem_desc(sbase + C_::OFF_AM + b * 2 * AM_BYTES, 16, 1024, LAYOUT_SW128);
#pragma unroll
        for (int j = 0; j < 3; ++j) {  // K = 24 joints = 3 steps of 8
          const uint64_t mh = umma_desc_advance(mh0, j * 32), ml = umma_desc_advance(mh0, AM_BYTES + j * 32);
          umma_tf32_ts(d, wl + j * 8, mh, idesc, j > 0 ? 1u : 0u);
          umma_tf32_ts(d, wh + j * 8, ml, idesc, 1u);
          umma_tf32_ts(d, wh + j * 8, mh, idesc, 1u);
        }
        umma_commit(tfull);
        umma_commit(amfree(b));
      }
    }
    __syncwarp();
  } else if (warp == W_ISSUE) {
    // ------------------------------------------------------------------ dA GEMM issuer
    if (lane == 0) {
      // LBO = stride between n blocks, SBO = 512 (4 k rows); a k step of 8 vertices = 1024 bytes
      const uint64_t gh0 = umma_smem_desc(sbase + C_::OFF_G, C_::GQ_BYTES, 512, LAYOUT_SW128_BASE32B);
      const uint32_t wt = tmem_base + C_::TM_WT;
      tprev = (uint32_t)clock();
      // one MMA per k step: N = 192 = [G_hi | G_lo] (a tf32 MMA costs ~96 cycles whatever its N up to ~176 and
      // ~106 at N = 192 -- measured, tools/ubench/mma_pace.cu -- so the kernel is paced by its MMA COUNT)
      for (int gi = 0; gi < ng; ++gi) {
        const int b = gi & 1;
        const uint32_t use = gi >> 1;
        const uint32_t boff = (uint32_t)(16 * b) << 16;  // accumulator AND operand copy of buffer b
        const uint32_t d = tmem_base + C_::TM_D + boff;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          tick(4 + kb);
          mbar_wait(gfull(kb), gi & 1);
          tick(kb);
          if (kb == 0 && use > 0) mbar_wait(d2empty(b), (use - 1) & 1);
          tc_fence_after();
          if (!(a.ablate & 4))
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32_ts(d, wt + boff + kb * 32 + ks * 8, umma_desc_advance(gh0, (kb * 32 + ks * 8) * 128), idesc_d,
                         (kb | ks) ? 1u : 0u);
          umma_commit(gfree(kb));
        }
        umma_commit(d2full(b));
      }
      tick_flush(2);
    }
    __syncwarp();
  } else if (warp >= W_DRAIN && warp < W_DRAIN + 8) {
    // ------------------------------------------------------------------ accumulator drain: one warp per lane quarter
    // Buffer b of the accumulator lives in lanes [16b, 16b + 16) of every quarter; quarter q carries joints 6q..6q+5:
    // lanes +0..5 = W_hi^T rows (columns [. G_hi | . G_lo]), lanes +6..11 = W_lo^T rows (only . G_hi counts).
    // dA[j][f][.] = hi[. G_hi] + hi[. G_lo] + lo[. G_hi]: two column runs of the own lane plus one shuffle from six
    // lanes up, all in registers; the sums go straight into the group's slab (laid out like global dA) and ONE thread
    // hands the 9 KB slab to the bulk-reduce engine (cp.reduce.async.bulk .add.f32): no per-thread atomics, the LSU
    // never sees the 54-way accumulation over vertex tiles.  With two accumulators and two slabs all of this overlaps
    // the next group's MMAs.
    // two warps per quarter: frame pairs {0, 1} and {2, 3}
    const int q = (warp - W_DRAIN) & 3, fh = (warp - W_DRAIN) >> 2, dt = (warp - W_DRAIN) * 32 + lane;
    const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + C_::TM_D;
    tprev = (uint32_t)clock();
    for (int gi = 0; gi < ng; ++gi) {
      const int b = gi & 1;
      mbar_wait(d2full(b), (gi >> 1) & 1);
      tick(0);
      tc_fence_after();
      const uint32_t slab = sbase + C_::OFF_STG + b * C_::STG_BYTES;
      if (gi >= 2) {  // the bulk reduction that read this slab two groups ago has left shared memory
        if (dt == 0) bulk_wait_read1();
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
      tick(1);
      const int jl = lane - 16 * b;          // 0..5: W_hi row of joint 6q + jl
      const bool on = jl >= 0 && jl < 6;
      const uint32_t srow = slab + (6 * q + jl) * 48;
#pragma unroll 1
      for (int fp = 2 * fh; fp < 2 * fh + 2; ++fp) {  // two frames = 24 columns of each half per batch
        uint32_t t[24], u[24];
        tmem_ld16(ta + fp * 24, t);
        tmem_ld8(ta + fp * 24 + 16, t + 16);
        tmem_ld16(ta + NROW + fp * 24, u);
        tmem_ld8(ta + NROW + fp * 24 + 16, u + 16);
        tmem_ld_wait();
        float v[24];
#pragma unroll
        for (int i = 0; i < 24; ++i)
          v[i] = __uint_as_float(t[i]) + __uint_as_float(u[i]) + __shfl_down_sync(0xffffffffu, __uint_as_float(t[i]), 6);
        if (on) {
#pragma unroll
          for (int m = 0; m < 6; ++m)
            sts128(srow + (2 * fp + m / 3) * 1152 + (m % 3) * 16, make_float4(v[4 * m], v[4 * m + 1], v[4 * m + 2], v[4 * m + 3]));
        }
      }
      tc_fence_before();
      mbar_arrive(d2empty(b));
      tick(2);
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, 256;" ::: "memory");  // slab complete
      if (dt == 0 && !(a.ablate & 1)) {
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
    for (int gi = 0; gi < ng; ++gi) {
      const int st = gi % C_::NSTAGE;
      const uint32_t use = gi / C_::NSTAGE;
      if (use > 0) mbar_wait(sempty(st), (use - 1) & 1);
      const uint32_t sa = sbase + C_::OFF_ST + st * C_::ST_BYTES;
      const int n0 = group_of(gi) * FB;
      if (a.ablate & 16) {
        if (lane == 0) mbar_arrive(sfull(st));
      } else {
        // lanes 0..23: pose offsets of (frame, component) = (lane / 3, lane % 3); lanes 24..31: dV of frame lane - 24
        const void* src;
        uint32_t dst, bytes;
        if (lane < 24) {
          src = a.PO + (size_t)min(n0 + lane / 3, last) * (3 * VP) + (lane % 3) * VP + v0;
          dst = sa + lane * (TILE_V * 4);
          bytes = TILE_V * 4;
        } else {
          const int f = lane - 24, n = min(n0 + f, last), odd = n & 1;
          src = reinterpret_cast<const char*>(a.dV + (size_t)n * (NV * 3)) + v0 * 12 - 8 * odd;
          dst = sa + C_::ST_PO_BYTES + f * C_::DV_SLOT;
          // last tile: the frame ends after 106 vertices (1272 bytes); an even frame is cut to a 16-byte multiple
          // and the worker of the last vertex fetches its y, z itself
          bytes = vt == NT_V - 1 ? (odd ? 1280u : 1264u) : (odd ? TILE_V * 12 + 16 : TILE_V * 12);
          if (!HAS_DV) bytes = 0;
        }
        const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
        if (lane == 0) mbar_arrive_expect_tx(sfull(st), total);
        __syncwarp();
        if (bytes) bulk_g2s(dst, src, bytes, sfull(st));
      }
    }
  } else if (warp < NWORK) {
    // ------------------------------------------------------------------ workers: four threads per vertex
    // warp = 4 * (frame pair) + (TMEM lane quarter): the four warps of a quarter share one warp scheduler and one
    // G k block and take two frames of the group each, so their barrier / TMEM latencies overlap
    constexpr int FH = FB * 4 / NWORK;  // frames of the group per worker warp
    static_assert(FH == 2, "two frames per worker warp");
    const int wq = warp & 3, half = warp >> 2;
    const int v = v0 + wq * 32 + lane;
    const bool real = v < NV;
    float bs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) bs[c] = __ldg(a.base + c * VP + v);
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    // G addresses.  16-byte unit u = 3 * (frame slot) + r holds [dv_r p0, dv_r p1, dv_r p2, dv_r]: n block u >> 3,
    // 32-byte chunk (u >> 1) & 3 stored at chunk position ^ (k & 3), half u & 1.  This thread owns units 6*half ..
    // 6*half+5 = three chunks; within a chunk the lanes with bit 2 set store the odd unit first, so the eight lanes
    // of a store phase cover all eight 16-byte slots of the 128-byte bank window (conflict free).  One 16-byte store
    // per unit and part: a quarter of the LSU instructions of a K-major G (the LSU instruction rate, ~0.5 / cycle, is
    // what the workers run into once the accumulator is double buffered).
    const bool swp = (lane >> 2) & 1;
    uint32_t gx[3], gy[3];
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      const int p = 3 * half + pl;
      const uint32_t cb = sbase + C_::OFF_G + (p >> 2) * C_::GQ_BYTES + (wq * 32 + lane) * 128 +
                          ((((p & 3) ^ lane) & 3) << 5);
      gx[pl] = cb + (swp ? 16u : 0u);
      gy[pl] = cb + (swp ? 0u : 16u);
    }
    auto put = [&](uint32_t addr, float d, float p0, float p1, float p2) {
      const float4 g = make_float4(d * p0, d * p1, d * p2, d);
      sts128(addr, g);
      sts128(addr + C_::GT_LO,
             make_float4(g.x - trunc_tf32(g.x), g.y - trunc_tf32(g.y), g.z - trunc_tf32(g.z), g.w - trunc_tf32(g.w)));
    };
    const unsigned lbit = 1u << lane;
    const float sp = a.sign_scale, sn = -a.sign_scale;

    uint32_t sA = 0u, sB = 0u;  // this half's 12 sign words (two frames), one per lane, prefetched one group ahead
    auto load_signs = [&](int gi, uint32_t& w) {
      if (!HAS_SIGNS) return;
      const int n0 = group_of(gi < ng ? gi : ng - 1) * FB + half * FH;
      const int fc = lane >> 1;  // word (f*3 + c)*2 + {pos, neg}
      const int n = min(n0 + fc / 3, last);
      w = lane < FH * 6 ? __ldg(a.signs + (((size_t)n * 3 + fc % 3) * 2 + (lane & 1)) * (VP / 32) + (v >> 5)) : 0u;
    };
    const uint32_t st_po = sbase + C_::OFF_ST + (half * FH * 3) * (TILE_V * 4) + (wq * 32 + lane) * 4;
    const uint32_t st_dv = sbase + C_::OFF_ST + C_::ST_PO_BYTES + (half * FH) * C_::DV_SLOT + (wq * 32 + lane) * 12;
    auto body = [&](int gi, uint32_t sw, uint32_t& swn) {
      const int n0 = group_of(gi) * FB + half * FH;
      load_signs(gi + 1, swn);
      // ---- this group's inputs from the ring
      float po[FH][3], dv[FH][3];
      {
        const int st = gi % C_::NSTAGE;
        mbar_wait(sfull(st), (gi / C_::NSTAGE) & 1);
        const uint32_t so = st * C_::ST_BYTES;
#pragma unroll
        for (int f = 0; f < FH; ++f)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            po[f][c] = lds32(st_po + so + (f * 3 + c) * (TILE_V * 4));
            if (HAS_DV) {
              const int n = min(n0 + f, last);
              float x = lds32(st_dv + so + f * C_::DV_SLOT + 8 * (n & 1) + c * 4);
              if (v == NV - 1 && !(n & 1) && c > 0) x = __ldg(a.dV + ((size_t)n * NV + v) * 3 + c);  // see the loader
              dv[f][c] = real ? x : 0.f;
            } else {
              dv[f][c] = 0.f;
            }
          }
        if (TIMING) {  // force the arrival
          float acc = 0.f;
          for (int f = 0; f < FH; ++f)
            for (int c = 0; c < 3; ++c) acc += po[f][c] + dv[f][c];
          if (acc == 1.2345e-30f) tk[7] += 1;
        }
        mbar_arrive(sempty(st));
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
      // ---- dPO = R^T dv from the T rows in TMEM
      tick(1);
      mbar_wait(tfull, gi & 1);
      tick(2);
      tc_fence_after();
      uint32_t t[FH * 12];
      const uint32_t ta = tmem_base + lane_addr + C_::TM_T + half * (FH * 12);
      tmem_ld16(ta, t);
      tmem_ld8(ta + 16, t + 16);
      tmem_ld_wait();
      tick(3);
      tc_fence_before();
      mbar_arrive(tempty);
      {
        float* d = a.dPO + (size_t)n0 * (3 * VP) + v;
#pragma unroll
        for (int f = 0; f < FH; ++f) {
          if (n0 + f <= last && !(a.ablate & 8)) {  // warp-uniform
#pragma unroll
            for (int c = 0; c < 3; ++c)
              __stcs(d + (size_t)f * (3 * VP) + c * VP,
                     fmaf(__uint_as_float(t[f * 12 + c]), dv[f][0],
                          fmaf(__uint_as_float(t[f * 12 + 4 + c]), dv[f][1], __uint_as_float(t[f * 12 + 8 + c]) * dv[f][2])));
          }
        }
      }
      // ---- G for the dA GEMM (the previous group's MMAs over this k block must have finished reading it)
      tick(4);
      if (gi > 0) mbar_wait(gfree(wq), (gi - 1) & 1);
      tick(5);
      if (!(a.ablate & 2)) {
        const float p0[3] = {bs[0] + po[0][0], bs[1] + po[0][1], bs[2] + po[0][2]};
        const float p1[3] = {bs[0] + po[1][0], bs[1] + po[1][1], bs[2] + po[1][2]};
        // chunk 0: units (f0,r0) | (f0,r1); chunk 1: (f0,r2) | (f1,r0); chunk 2: (f1,r1) | (f1,r2)
        put(gx[0], swp ? dv[0][1] : dv[0][0], p0[0], p0[1], p0[2]);
        put(gy[0], swp ? dv[0][0] : dv[0][1], p0[0], p0[1], p0[2]);
        put(gx[1], swp ? dv[1][0] : dv[0][2], swp ? p1[0] : p0[0], swp ? p1[1] : p0[1], swp ? p1[2] : p0[2]);
        put(gy[1], swp ? dv[0][2] : dv[1][0], swp ? p0[0] : p1[0], swp ? p0[1] : p1[1], swp ? p0[2] : p1[2]);
        put(gx[2], swp ? dv[1][2] : dv[1][1], p1[0], p1[1], p1[2]);
        put(gy[2], swp ? dv[1][1] : dv[1][2], p1[0], p1[1], p1[2]);
      }
      fence_proxy_async_smem();
      mbar_arrive(gfull(wq));
      tick(6);
    };
    load_signs(0, sA);
    tprev = (uint32_t)clock();
    for (int gi = 0; gi < ng; gi += 2) {
      body(gi, sA, sB);
      if (gi + 1 < ng) body(gi + 1, sB, sA);
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

}  // namespace

namespace nb {
int skin_fwd2(const float* A, const float* PO, const float* base, const float* Wt, float* V, int N, cudaStream_t s);
int skin_l1_fwd2(const float* A_a, const float* PO_a, const float* A_b, const float* PO_b, const float* base,
                 const float* Wt, uint32_t* signs, float* l1_sum, int N, cudaStream_t s);
void skin_chunking(int N, int frames_per_group, int ctas_per_sm, double setup, int max_chunks, int& groups_per_chunk,
                   int& nchunks);
int skin_bwd2(const float* A, const float* PO, const float* base, const float* Wt, const float* dV,
              const uint32_t* signs, float sign_scale, float* dPO, float* dA, int N, cudaStream_t s);
}  // namespace nb

namespace {
static int bwd_generation() {
  static const int g = [] { const char* e = getenv("NEMO_B200_SKIN_BWD"); return e ? atoi(e) : 2; }();
  return g;
}
static int fwd_generation() {
  static const int g = [] { const char* e = getenv("NEMO_B200_SKIN_FWD"); return e ? atoi(e) : 2; }();
  return g;
}

// Frame groups per CTA.  Every CTA pays a set-up (W operands into shared / tensor memory) worth `setup` groups, and
// the grid runs in waves over 148 x ctas_per_sm slots: pick the chunk count that minimises waves x (groups + set-up).
void chunking(int N, int frames_per_group, int ctas_per_sm, double setup, int max_chunks, int& groups_per_chunk,
              int& nchunks) {
  const int groups = cdiv(N, frames_per_group);
  static const int forced = [] { const char* e = getenv("NEMO_B200_SKIN_CHUNKS"); return e ? max(0, atoi(e)) : 0; }();
  int best = 1;
  if (forced) {
    best = min(forced, groups);
  } else {
    double best_cost = 1e30;
    for (int c = 1; c <= max_chunks && c <= groups; ++c) {
      const int gpc = cdiv(groups, c), ctas = NT_V * cdiv(groups, gpc);
      const double cost = (double)cdiv(ctas, num_sms() * ctas_per_sm) * (gpc + setup);
      if (cost < best_cost) best_cost = cost, best = c;
    }
  }
  groups_per_chunk = cdiv(groups, best);
  nchunks = cdiv(groups, groups_per_chunk);
}
}  // namespace
void nb::skin_chunking(int N, int frames_per_group, int ctas_per_sm, double setup, int max_chunks, int& groups_per_chunk,
                       int& nchunks) {
  chunking(N, frames_per_group, ctas_per_sm, setup, max_chunks, groups_per_chunk, nchunks);
}
namespace {

template <int MODE>
int launch_fwd(const SkinArgs& a, cudaStream_t s) {
  NB_SET_SMEM_ONCE(k_skin_tc_fwd<MODE>, FwdCfg::SMEM_BYTES);
  k_skin_tc_fwd<MODE><<<NT_V * a.nchunks, FWD_THREADS, FwdCfg::SMEM_BYTES, s>>>(a);
  NB_LAUNCH_CHECK("skin_tc_fwd");
}

template <bool HAS_DV, bool HAS_SIGNS>
int launch_bwd_plain(const SkinBwdArgs& a, cudaStream_t s) {
  NB_SET_SMEM_ONCE((k_skin_tc_bwd<HAS_DV, HAS_SIGNS>), BwdCfg::SMEM_BYTES);
  k_skin_tc_bwd<HAS_DV, HAS_SIGNS><<<NT_V * a.nchunks, BWD_THREADS, BwdCfg::SMEM_BYTES, s>>>(a);
  NB_LAUNCH_CHECK("skin_tc_bwd");
}

// NEMO_B200_SKIN_TIMING=1 (debug): run the instrumented dV kernel, synchronise and print the phase counters
int launch_bwd_timing(SkinBwdArgs a, cudaStream_t s) {
  static unsigned long long* dbg = nullptr;
  if (!dbg) NB_CUDA(cudaMalloc(&dbg, 40 * sizeof(unsigned long long)));
  NB_CUDA(cudaMemsetAsync(dbg, 0, 40 * sizeof(unsigned long long), s));
  a.dbg = dbg;
  NB_CUDA(cudaFuncSetAttribute(k_skin_tc_bwd<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               BwdCfg::SMEM_BYTES));
  k_skin_tc_bwd<true, false, true><<<NT_V * a.nchunks, BWD_THREADS, BwdCfg::SMEM_BYTES, s>>>(a);
  NB_CUDA(cudaStreamSynchronize(s));
  unsigned long long h[40];
  NB_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
  const double groups = (double)NT_V * cdiv(a.N, FB);
  static const char* names[5][8] = {
      {"inputs (ring)", "signs", "wait tfull", "tmem ld", "dPO", "wait gfree", "G phase", "-"},
      {"wait amfree", "Amat store", "fetch", "-", "-", "-", "-", "-"},
      {"wait gfull0", "wait gfull1", "wait gfull2", "wait gfull3", "mma0", "mma1", "mma2", "mma3"},
      {"wait d2full", "wait slab free", "copy + sum", "bulk reduce", "-", "-", "-", "-"},
      {"-", "-", "-", "-", "-", "-", "-", "-"}};
  const double per[5] = {(double)NWORK, 1.0, 1.0, 1.0, 1.0};  // warps per role
  for (int r = 0; r < 5; ++r) {
    fprintf(stderr, "[skin_bwd timing] role %d:", r);
    for (int i = 0; i < 8; ++i)
      if (names[r][i][0] != '-') fprintf(stderr, "  %s %.0f", names[r][i], (double)h[r * 8 + i] / groups / per[r]);
    fprintf(stderr, "  (cycles per group per warp)\n");
  }
  return 0;
}

template <bool HAS_DV, bool HAS_SIGNS>
int launch_bwd(const SkinBwdArgs& a, cudaStream_t s) {
  static const int timing = [] { const char* e = getenv("NEMO_B200_SKIN_TIMING"); return e ? atoi(e) : 0; }();
  if (timing && HAS_DV && !HAS_SIGNS) return launch_bwd_timing(a, s);
  return launch_bwd_plain<HAS_DV, HAS_SIGNS>(a, s);
}

}  // namespace

extern "C" int nemo_b200_skin_fwd(const float* A, const float* PO, const float* base, const float* Wt, float* V, int N,
                                  void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(A && PO && base && Wt && V && N >= 0, "skin_fwd: bad args");
  if (fwd_generation() >= 2 && ((uintptr_t)PO & 15) == 0) return nb::skin_fwd2(A, PO, base, Wt, V, N, S(stream));
  SkinArgs a{};
  a.A = A; a.PO = PO; a.base = base; a.Wt = Wt; a.V = V; a.N = N;
  chunking(N, FB, 2, 2.0, 32, a.groups_per_chunk, a.nchunks);
  return launch_fwd<0>(a, S(stream));
}

extern "C" int nemo_b200_skin_l1_fwd(const float* A_a, const float* PO_a, const float* A_b, const float* PO_b,
                                     const float* base, const float* Wt, uint32_t* signs, float* l1_sum, int N,
                                     void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(A_a && PO_a && A_b && PO_b && base && Wt && signs && l1_sum && N >= 0, "skin_l1_fwd: bad args");
  if (fwd_generation() >= 2 && (((uintptr_t)PO_a | (uintptr_t)PO_b) & 15) == 0)
    return nb::skin_l1_fwd2(A_a, PO_a, A_b, PO_b, base, Wt, signs, l1_sum, N, S(stream));
  SkinArgs a{};
  a.A = A_a; a.PO = PO_a; a.A2 = A_b; a.PO2 = PO_b; a.base = base; a.Wt = Wt; a.signs = signs; a.l1_sum = l1_sum;
  a.N = N;
  chunking(N, FB / 2, 2, 2.0, 32, a.groups_per_chunk, a.nchunks);
  return launch_fwd<1>(a, S(stream));
}

extern "C" int nemo_b200_skin_bwd(const float* A, const float* PO, const float* base, const float* Wt, const float* dV,
                                  const uint32_t* signs, float sign_scale, float* dPO, float* dA, int N,
                                  void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(A && PO && base && Wt && dPO && dA && (dV || signs) && N >= 0, "skin_bwd: bad args");
  cudaStream_t s = S(stream);
  // second generation: TMA tensor loads (PO and dV 16-byte aligned, which every cudaMalloc'ed / torch buffer is)
  if (bwd_generation() >= 2 && ((uintptr_t)PO & 15) == 0 && ((uintptr_t)dV & 15) == 0)
    return nb::skin_bwd2(A, PO, base, Wt, dV, signs, sign_scale, dPO, dA, N, s);
  SkinBwdArgs a{};
  a.A = A; a.PO = PO; a.base = base; a.Wt = Wt; a.dV = dV; a.signs = signs;
  a.sign_scale = sign_scale; a.dPO = dPO; a.dA = dA; a.N = N;
  chunking(N, FB, 1, 2.0, 16, a.groups_per_chunk, a.nchunks);
  static const int abl = [] { const char* e = getenv("NEMO_B200_SKIN_ABLATE"); return e ? atoi(e) : 0; }();
  a.ablate = abl;
  NB_CUDA(cudaMemsetAsync(dA, 0, sizeof(float) * (size_t)N * 288, s));  // vertex tiles accumulate with reductions
  if (dV && signs) return launch_bwd<true, true>(a, s);
  if (dV) return launch_bwd<true, false>(a, s);
  return launch_bwd<false, true>(a, s);
}
