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
// 16-byte vector reduction (sm_90+): one L2 atomic transaction for four consecutive floats
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
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
  static constexpr int OFF_RED = OFF_AM + 4 * AM_BYTES;
  static constexpr int OFF_BARS = OFF_RED + 256;
  static constexpr int SMEM_BYTES = OFF_BARS + 64 + 1024;
  static constexpr int TMEM_COLS = 256;                 // T: 2 x 128
};

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
__global__ void __launch_bounds__(256, 2) k_skin_tc_fwd(const SkinArgs a) {
  using C_ = FwdCfg;
  constexpr int FG = MODE == 0 ? FB : FB / 2;  // frames per group
  constexpr int NMESH = MODE == 0 ? 1 : 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + C_::OFF_BARS;
  auto tfull = [&](int b) { return bars + 8u * b; };         // T of buffer b complete          (commit)
  auto tempty = [&](int b) { return bars + 8u * (2 + b); };  // workers done reading T[b]       (128)
  auto amfree = [&](int b) { return bars + 8u * (4 + b); };  // MMAs done reading Amat[b]       (commit)
  const uint32_t tmem_slot = bars + 48;
  const uint32_t sred = sbase + C_::OFF_RED;

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
    }
    fence_barrier_init();
  }
  if (warp == 7) tmem_alloc<C_::TMEM_COLS>(tmem_slot);
  // zero both B-operand buffers once: the k = 24..31 pad columns stay zero for the whole kernel
  for (int i = tid; i < 4 * AM_BYTES / 16; i += 256)
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
    // ------------------------------------------------------------------ producers (96 threads): Amat + MMA issue
    if (warp < 7) {
      const int pt = tid - 128;
      AmatWriter aw;
      aw.init(pt);
      for (int gi = 0; gi < ng; ++gi) {
        const int b = gi & 1;
        const uint32_t use = gi >> 1;  // how many times buffer b was used before
        if (use > 0) mbar_wait(amfree(b), (use - 1) & 1);
        const uint32_t am = sbase + C_::OFF_AM + b * 2 * AM_BYTES;
        const int n0 = group_of(gi) * FG;
        aw.write(am, pt, [&](int s) {
          const float* src = (NMESH == 2 && s >= FG) ? a.A2 : a.A;
          return src + (size_t)min(n0 + (NMESH == 2 ? (s & (FG - 1)) : s), last) * 288;
        });
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 96;" ::: "memory");
        if (pt == 0) {
          if (use > 0) mbar_wait(tempty(b), (use - 1) & 1);
          tc_fence_after();
          const uint32_t d = tmem_base + b * 128;
          const uint64_t wh0 = umma_smem_desc(sbase + C_::OFF_W, 16, 1024, LAYOUT_SW128);
          const uint64_t mh0 = umma_smem_desc(am, 16, 1024, LAYOUT_SW128);
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
    }
  } else {
    // ------------------------------------------------------------------ workers: one thread per vertex
    const int v = v0 + tid;
    const bool real = v < NV;
    float bs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) bs[c] = __ldg(a.base + c * VP + v);
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float l1 = 0.f;
    // pose offsets are prefetched one group ahead into the other register set: HBM reads overlap the arithmetic
    float pa[FB][3], pb[FB][3];
    auto load_po = [&](int gi, float (&dst)[FB][3]) {
      const int n0 = group_of(gi < ng ? gi : ng - 1) * FG;
#pragma unroll
      for (int s = 0; s < FB; ++s) {
        const float* P = (NMESH == 2 && s >= FG) ? a.PO2 : a.PO;
        const int n = min(n0 + (NMESH == 2 ? (s & (FG - 1)) : s), last);
        const float* p = P + (size_t)n * (3 * VP) + v;
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[s][c] = __ldcs(p + c * VP);
      }
    };
    auto body = [&](int gi, float (&po)[FB][3], float (&pn)[FB][3]) {
      const int b = gi & 1;
      const int n0 = group_of(gi) * FG;
      load_po(gi + 1, pn);
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
    load_po(0, pa);
    for (int gi = 0; gi < ng; gi += 2) {
      body(gi, pa, pb);
      if (gi + 1 < ng) body(gi + 1, pb, pa);
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
    tmem_dealloc<C_::TMEM_COLS>(tmem_base);
  }
}

// ================================================================================================ backward
// Per vertex tile and group of 8 frames, two GEMMs on tcgen05:
//   T[v, (f,12)]    = W[v, :] . Amat[:, (f,12)]      A operand (W tile, tf32 hi / lo) lives in TENSOR MEMORY, so only the
//                                                    3 KB B operand is fetched from shared memory per MMA
//   dA^T[(f,12), j] = sum_v G^T[(f,12), v] . W[v, j] G[v,(f,r,c)] = dv[f,v,r] * [v_posed;1][c]
// The worker thread of vertex v reads the rotation rows of T from its TMEM lane, assembles dv (upstream dV and / or
// the L1 sign bits), writes dPO = R^T dv (coalesced planar) and scatters G (raw fp32 = the hi operand, the tensor core
// ignores the low 13 mantissa bits) and lo = G - trunc(G) into the K-major A operand of the second GEMM (row =
// (f,r,c), k = vertex: a warp writes one full 128-byte swizzled row per store, conflict free).
// The B operand of the second GEMM is [W hi | W lo] side by side (64 rows), so 3xTF32 costs two MMAs per k step
// instead of three and G^T is fetched twice instead of three times:
//   D[:, 0:32] += G_hi.W_hi + G_lo.W_hi      D[:, 32:64] += G_hi.W_lo        dA^T = D[:, j] + D[:, 32+j]
// G^T is synchronised per 32-vertex k block (the two worker warps of a TMEM lane quarter share one warp scheduler
// and one k block), so the MMAs of one block overlap the G writes of the others.  Three producer warps drain the
// accumulator (lanes = (f,r,c), columns = joints) into global dAt (N,12,24) with 16-byte vector reductions
// (the 54 vertex tiles add up).
struct BwdCfg {
  static constexpr int WD_ROWS = 48;                 // [W hi (24 joints) | W lo (24 joints)]
  static constexpr int WD_KB_BYTES = WD_ROWS * 128;  // one 32-vertex k block of [W hi | W lo]: 48 rows(j) x 128 bytes
  static constexpr int GT_KB_BYTES = NROW * 128;     // one k block of one tf32 part of G^T: 96 rows x 128 bytes
  static constexpr int GT_BYTES = 4 * GT_KB_BYTES;   // rows 96..127 read by an M = 128 MMA alias the next block (unused)
  static constexpr int OFF_WD = 0;                               // 4 x 6 KB
  static constexpr int OFF_AM = OFF_WD + 4 * WD_KB_BYTES;        // 2 buffers x [hi][lo] x 12 KB
  static constexpr int OFF_GT = OFF_AM + 4 * AM_BYTES;           // [hi][lo] x 48 KB
  static constexpr int OFF_TAIL = OFF_GT + 2 * GT_BYTES;         // 4 KB of zeros read by the last block's rows 96..127
  static constexpr int OFF_BARS = OFF_TAIL + 4096;
  static constexpr int SMEM_BYTES = OFF_BARS + 256 + 1024;
  static constexpr int TM_T = 0;        // T: 2 x 128 columns
  static constexpr int TM_D = 256;      // dA^T: 2 x 64 columns (48 used: [.W hi | .W lo])
  static constexpr int TM_W = 384;      // W tile: 32 hi + 32 lo columns
  static constexpr int TMEM_COLS = 512;
};
static_assert(BwdCfg::SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct SkinBwdArgs {
  const float* A; const float* PO; const float* base; const float* Wt;
  const float* dV; const uint32_t* signs;
  float* dPO; float* dAt;
  float sign_scale;
  int N, groups_per_chunk, nchunks;
  unsigned long long* dbg;  // TIMING instantiation only: per-role phase cycle counters
  int ablate;  // debug (NEMO_B200_SKIN_ABLATE): 1 no dA reductions, 2 no G stores, 4 no dA MMAs, 8 no dPO stores, 16 no loads
};

// NWORK worker warps (8 or 16): warp = 4 * (frame slice of the group) + (TMEM lane quarter); 3 producer warps; 1 issuer
template <bool HAS_DV, bool HAS_SIGNS, int NWORK, bool TIMING = false>
__global__ void __launch_bounds__((NWORK + 4) * 32, 1) k_skin_tc_bwd(const SkinBwdArgs a) {
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
  constexpr int THREADS = (NWORK + 4) * 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + C_::OFF_BARS;
  auto tfull = [&](int b) { return bars + 8u * b; };            // T[b] complete                     (commit)
  auto tempty = [&](int b) { return bars + 8u * (2 + b); };     // workers done with T[b]            (512)
  auto amfree = [&](int b) { return bars + 8u * (4 + b); };     // T GEMM done reading Amat[b]       (commit)
  auto d2full = [&](int b) { return bars + 8u * (6 + b); };     // dA^T[b] complete                  (commit)
  auto d2empty = [&](int b) { return bars + 8u * (8 + b); };    // dA^T[b] drained                   (96)
  auto gfull = [&](int kb) { return bars + 8u * (10 + kb); };   // G^T k block written               (128)
  auto gfree = [&](int kb) { return bars + 8u * (14 + kb); };   // dA GEMM done reading the k block  (commit)
  const uint32_t tmem_slot = bars + 8u * 18;

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
      mbar_init(tempty(b), NWORK * 32);
      mbar_init(amfree(b), 1);
      mbar_init(d2full(b), 1);
      mbar_init(d2empty(b), 96);
    }
    for (int kb = 0; kb < 4; ++kb) {
      mbar_init(gfull(kb), NWORK * 8);
      mbar_init(gfree(kb), 1);
    }
    fence_barrier_init();
  }
  constexpr int W_PROD = NWORK, W_ISSUE = NWORK + 3;  // warps [W_PROD, W_PROD+3): producers; W_ISSUE: dA GEMM
  if (warp == W_ISSUE) tmem_alloc<C_::TMEM_COLS>(tmem_slot);
  // zero Amat (pad columns k = 24..31), G^T and its tail
  for (int i = tid; i < (C_::OFF_BARS - C_::OFF_AM) / 16; i += THREADS)
    sts128(sbase + C_::OFF_AM + 16 * i, make_float4(0.f, 0.f, 0.f, 0.f));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32u(tmem_slot);
  if (tid < TILE_V) {
    // this vertex's 24 weights: tf32 hi / lo into its TMEM lane (A operand of the T GEMM) and into the
    // [W hi | W lo] B operand of the dA GEMM (row = joint, k = vertex; k block = warp)
    const int v = v0 + tid;
    uint32_t h[32], l[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      float w = k < NJ ? __ldg(a.Wt + (size_t)k * VP + v) : 0.f, hi, lo;
      tf32_split(w, hi, lo);
      h[k] = __float_as_uint(hi);
      l[k] = __float_as_uint(lo);
      const uint32_t o = sbase + C_::OFF_WD + (uint32_t)(warp * C_::WD_KB_BYTES);
      if (k < NJ) {
        sts32(o + kmajor_sw128(k, lane), hi);
        sts32(o + kmajor_sw128(NJ + k, lane), lo);
      }
    }
    const uint32_t tw = tmem_base + ((uint32_t)(warp * 32) << 16) + C_::TM_W;
    tmem_st32(tw, h);
    tmem_st32(tw + 32, l);
    tmem_st_wait();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t idesc_t = umma_idesc_tf32(128, NROW, 0, 0);
  constexpr uint32_t idesc_d48 = umma_idesc_tf32(128, C_::WD_ROWS, 0, 0);
  constexpr uint32_t idesc_d32 = umma_idesc_tf32(128, 32, 0, 0);

  if (warp >= W_PROD && warp < W_ISSUE) {
    // ------------------------------------------------------------------ producers: Amat, T GEMM issue, dA^T drain
    const int pt = tid - W_PROD * 32;
    AmatWriter aw;
    aw.init(pt);
    const int row = (warp & 3) * 32 + lane;  // accumulator lane = (f, r, c); W_PROD % 4 == 0
    const int rf = row / 12, rq = row % 12;
    auto drain = [&](int gi) {
      const int b = gi & 1;
      mbar_wait(d2full(b), (gi >> 1) & 1);
      tc_fence_after();
      uint32_t t[24], u[24];
      const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C_::TM_D + b * 64;
      tmem_ld16(ta, t);
      tmem_ld8(ta + 16, t + 16);
      tmem_ld16(ta + NJ, u);
      tmem_ld8(ta + NJ + 16, u + 16);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(d2empty(b));
      const int n = group_of(gi) * FB + rf;
      if (n <= last && !(a.ablate & 1)) {
        float* d = a.dAt + (size_t)n * 288 + rq * 24;
#pragma unroll
        for (int j = 0; j < NJ; j += 4)
          red_add_v4(d + j, __uint_as_float(t[j]) + __uint_as_float(u[j]),
                     __uint_as_float(t[j + 1]) + __uint_as_float(u[j + 1]),
                     __uint_as_float(t[j + 2]) + __uint_as_float(u[j + 2]),
                     __uint_as_float(t[j + 3]) + __uint_as_float(u[j + 3]));
      }
    };
    // The transforms of group gi + 2 are fetched into registers and the T GEMM of group gi + 1 is issued BEFORE this
    // thread blocks in the drain of group gi: the workers never wait for T.
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
    auto issue_t = [&](int gi) {
      const int b = gi & 1;
      const uint32_t use = gi >> 1;
      if (use > 0) mbar_wait(amfree(b), (use - 1) & 1);
      const uint32_t am = sbase + C_::OFF_AM + b * 2 * AM_BYTES;
      aw.store(am, x);
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 96;" ::: "memory");
      if (pt == 0) {
        if (use > 0) mbar_wait(tempty(b), (use - 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + C_::TM_T + b * 128;
        const uint32_t wh = tmem_base + C_::TM_W, wl = wh + 32;
        const uint64_t mh0 = umma_smem_desc(am, 16, 1024, LAYOUT_SW128);
#pragma unroll
        for (int j = 0; j < 3; ++j) {  // K = 24 joints = 3 steps of 8
          const uint64_t mh = umma_desc_advance(mh0, j * 32), ml = umma_desc_advance(mh0, AM_BYTES + j * 32);
          umma_tf32_ts(d, wl + j * 8, mh, idesc_t, j > 0 ? 1u : 0u);
          umma_tf32_ts(d, wh + j * 8, ml, idesc_t, 1u);
          umma_tf32_ts(d, wh + j * 8, mh, idesc_t, 1u);
        }
        umma_commit(tfull(b));
        umma_commit(amfree(b));
      }
    };
    fetch(0);
    issue_t(0);
    if (ng > 1) fetch(1);
    tprev = (uint32_t)clock();
    for (int gi = 0; gi < ng; ++gi) {
      if (gi + 1 < ng) {
        issue_t(gi + 1);
        tick(0);
        if (gi + 2 < ng) fetch(gi + 2);
      }
      drain(gi);
      tick(1);
    }
    tick_flush(1);
  } else if (warp == W_ISSUE) {
    // ------------------------------------------------------------------ dA GEMM issuer
    if (lane == 0) {
      const uint64_t gh0 = umma_smem_desc(sbase + C_::OFF_GT, 16, 1024, LAYOUT_SW128);
      const uint64_t wd0 = umma_smem_desc(sbase + C_::OFF_WD, 16, 1024, LAYOUT_SW128);
      tprev = (uint32_t)clock();
      for (int gi = 0; gi < ng; ++gi) {
        const int b = gi & 1;
        const uint32_t use = gi >> 1;
        const uint32_t d = tmem_base + C_::TM_D + b * 64;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          tick(4 + kb);
          mbar_wait(gfull(kb), gi & 1);
          tick(kb);
          if (kb == 0 && use > 0) mbar_wait(d2empty(b), (use - 1) & 1);
          tc_fence_after();
          if (!(a.ablate & 4))
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t go = kb * C_::GT_KB_BYTES + ks * 32, wo = kb * C_::WD_KB_BYTES + ks * 32;
            const uint64_t gh = umma_desc_advance(gh0, go), gl = umma_desc_advance(gh0, C_::GT_BYTES + go);
            const uint64_t wd = umma_desc_advance(wd0, wo);
            umma_tf32(d, gh, wd, idesc_d48, (kb | ks) ? 1u : 0u);   // [G_hi.W_hi | G_hi.W_lo]
            // += G_lo.W_hi (N = 32: columns 24..31 also receive G_lo.W_lo of joints 0..7, a valid 4th-order term)
            umma_tf32(d, gl, wd, idesc_d32, 1u);
          }
          umma_commit(gfree(kb));
        }
        umma_commit(d2full(b));
      }
      tick_flush(2);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ workers: four threads per vertex
    // warp = 4 * (frame pair) + (TMEM lane quarter): the four warps of a quarter share one warp scheduler and one
    // G^T k block and take two frames of the group each, so their barrier / TMEM / HBM latencies overlap
    constexpr int FH = FB * 4 / NWORK;  // frames of the group per worker warp
    const int wq = warp & 3, half = warp >> 2;
    const int v = v0 + wq * 32 + lane;
    const bool real = v < NV;
    float bs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) bs[c] = __ldg(a.base + c * VP + v);
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    // G^T addresses: row = half*48 + f*12 + q (q = r*4 + c), k = lane.  The swizzle phase row & 7 = ((f&1)*4 + q) & 7
    // is a compile-time constant per (f, q): eight lane-dependent offsets cover every store
    const uint32_t gt = sbase + C_::OFF_GT + wq * C_::GT_KB_BYTES + half * (FH * 12 * 128) + ((lane & 3) << 2);
    uint32_t xo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xo[i] = gt + ((((lane >> 2) ^ i) & 7) << 4);
    const unsigned lbit = 1u << lane;
    const float sp = a.sign_scale, sn = -a.sign_scale;

    float pA[FH][3], dA_[FH][3], pB[FH][3], dB_[FH][3];
    uint32_t sA = 0u, sB = 0u;  // this half's 24 sign words, one per lane
    auto load_in = [&](int gi, float (&p)[FH][3], float (&d)[FH][3], uint32_t& w) {
      const int n0 = group_of(gi < ng ? gi : ng - 1) * FB + half * FH;
      if (a.ablate & 16) return;
      if (HAS_SIGNS) {
        const int fc = lane >> 1;  // word (f*3 + c)*2 + {pos, neg}
        const int n = min(n0 + fc / 3, last);
        w = lane < FH * 6 ? __ldg(a.signs + (((size_t)n * 3 + fc % 3) * 2 + (lane & 1)) * (VP / 32) + (v >> 5)) : 0u;
      }
#pragma unroll
      for (int f = 0; f < FH; ++f) {
        const int n = min(n0 + f, last);
        const float* pp = a.PO + (size_t)n * (3 * VP) + v;
#pragma unroll
        for (int c = 0; c < 3; ++c) p[f][c] = __ldcs(pp + c * VP);
        if (HAS_DV) {
          const float* dp = a.dV + ((size_t)n * NV + (real ? v : NV - 1)) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float x = __ldcs(dp + c);
            d[f][c] = real ? x : 0.f;
          }
        }
      }
    };
    auto body = [&](int gi, float (&po)[FH][3], float (&dv)[FH][3], uint32_t sw, float (&pn)[FH][3],
                    float (&dn)[FH][3], uint32_t& swn) {
      const int b = gi & 1;
      const int n0 = group_of(gi) * FB + half * FH;
      if (TIMING) {  // force the arrival of this group's inputs (prefetched during the previous group)
        float acc = 0.f;
        for (int f = 0; f < FH; ++f)
          for (int c = 0; c < 3; ++c) acc += po[f][c] + (HAS_DV ? dv[f][c] : 0.f);
        if (acc == 1.2345e-30f) tk[7] += 1;
        tick(0);
      }
      load_in(gi + 1, pn, dn, swn);
      // ---- upstream gradient of the vertex in each frame
      if (HAS_SIGNS) {
#pragma unroll
        for (int f = 0; f < FH; ++f)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const uint32_t pm = __shfl_sync(0xffffffffu, sw, (f * 3 + c) * 2);
            const uint32_t nm = __shfl_sync(0xffffffffu, sw, (f * 3 + c) * 2 + 1);
            const float s = (pm & lbit) ? sp : ((nm & lbit) ? sn : 0.f);
            dv[f][c] = HAS_DV ? dv[f][c] + s : s;
          }
      }
      // ---- dPO = R^T dv from the T rows in TMEM (all of this half's rows in one batch of loads)
      tick(1);
      mbar_wait(tfull(b), (gi >> 1) & 1);
      tick(2);
      tc_fence_after();
      uint32_t t[FH * 12];
      const uint32_t ta = tmem_base + lane_addr + C_::TM_T + b * 128 + half * (FH * 12);
      static_assert(FH == 2 || FH == 4, "TMEM loads below fetch 24 or 48 columns");
      tmem_ld16(ta, t);
      if (FH == 2) {
        tmem_ld8(ta + 16, t + 16);
      } else {
        tmem_ld16(ta + 16, t + 16);
        tmem_ld16(ta + 32, t + 32);
      }
      tmem_ld_wait();
      tick(3);
      tc_fence_before();
      mbar_arrive(tempty(b));
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
      // ---- G^T for the dA GEMM (the previous group's MMAs over this k block must have finished reading it)
      tick(4);
      if (gi > 0) mbar_wait(gfree(wq), (gi - 1) & 1);
      tick(5);
      if (!(a.ablate & 2))
#pragma unroll
      for (int f = 0; f < FH; ++f) {
        const float ph[3] = {bs[0] + po[f][0], bs[1] + po[f][1], bs[2] + po[f][2]};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int q = r * 4 + c;
            const float g = c < 3 ? dv[f][r] * ph[c] : dv[f][r];
            const uint32_t off = xo[((f & 1) * 4 + q) & 7] + (f * 12 + q) * 128;
            sts32(off, g);
            sts32(off + C_::GT_BYTES, g - trunc_tf32(g));
          }
      }
      fence_proxy_async_smem();
      mbar_arrive(gfull(wq));
      tick(6);
    };
    load_in(0, pA, dA_, sA);
    tprev = (uint32_t)clock();
    for (int gi = 0; gi < ng; gi += 2) {
      body(gi, pA, dA_, sA, pB, dB_, sB);
      if (gi + 1 < ng) body(gi + 1, pB, dB_, sB, pA, dA_, sA);
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

int num_sms() {
  static int n = [] {
    int dev = 0, v = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

// frame groups per CTA: enough CTAs for >= ~8 waves over the SMs (load balance), at least 4 groups per CTA so
// the W-tile set-up amortises
void chunking(int N, int frames_per_group, int& groups_per_chunk, int& nchunks) {
  const int groups = cdiv(N, frames_per_group);
  static const int waves = [] { const char* e = getenv("NEMO_B200_SKIN_WAVES"); return e ? max(1, atoi(e)) : 8; }();
  int want = max(1, (waves * num_sms()) / NT_V);  // chunks so that 54 * chunks ~ `waves` waves
  groups_per_chunk = max(min(4, groups), cdiv(groups, want));
  nchunks = cdiv(groups, groups_per_chunk);
}

template <int MODE>
int launch_fwd(const SkinArgs& a, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    NB_CUDA(cudaFuncSetAttribute(k_skin_tc_fwd<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdCfg::SMEM_BYTES));
    configured = true;
  }
  k_skin_tc_fwd<MODE><<<NT_V * a.nchunks, 256, FwdCfg::SMEM_BYTES, s>>>(a);
  NB_LAUNCH_CHECK("skin_tc_fwd");
}

template <bool HAS_DV, bool HAS_SIGNS, int NWORK>
int launch_bwd_n(const SkinBwdArgs& a, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    NB_CUDA(cudaFuncSetAttribute(k_skin_tc_bwd<HAS_DV, HAS_SIGNS, NWORK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 BwdCfg::SMEM_BYTES));
    configured = true;
  }
  k_skin_tc_bwd<HAS_DV, HAS_SIGNS, NWORK><<<NT_V * a.nchunks, (NWORK + 4) * 32, BwdCfg::SMEM_BYTES, s>>>(a);
  NB_LAUNCH_CHECK("skin_tc_bwd");
}

// NEMO_B200_SKIN_TIMING=1 (debug): run the instrumented dV kernel, synchronise and print the phase counters
int launch_bwd_timing(SkinBwdArgs a, cudaStream_t s) {
  static unsigned long long* dbg = nullptr;
  if (!dbg) NB_CUDA(cudaMalloc(&dbg, 24 * sizeof(unsigned long long)));
  NB_CUDA(cudaMemsetAsync(dbg, 0, 24 * sizeof(unsigned long long), s));
  a.dbg = dbg;
  NB_CUDA(cudaFuncSetAttribute(k_skin_tc_bwd<true, false, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               BwdCfg::SMEM_BYTES));
  k_skin_tc_bwd<true, false, 8, true><<<NT_V * a.nchunks, 12 * 32, BwdCfg::SMEM_BYTES, s>>>(a);
  NB_CUDA(cudaStreamSynchronize(s));
  unsigned long long h[24];
  NB_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
  const double groups = (double)NT_V * cdiv(a.N, FB);
  static const char* names[3][8] = {
      {"inputs arrive", "issue prefetch+signs", "wait tfull", "tmem ld", "dPO", "wait gfree", "G phase", "-"},
      {"issue_t", "fetch+drain", "-", "-", "-", "-", "-", "-"},
      {"wait gfull0", "wait gfull1", "wait gfull2", "wait gfull3", "mma0", "mma1", "mma2", "mma3"}};
  const double per[3] = {8.0, 3.0, 1.0};  // warps per role
  for (int r = 0; r < 3; ++r) {
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
  static const int nwork = [] { const char* e = getenv("NEMO_B200_SKIN_BWD_WARPS"); return e ? atoi(e) : 16; }();
  return nwork == 8 ? launch_bwd_n<HAS_DV, HAS_SIGNS, 8>(a, s) : launch_bwd_n<HAS_DV, HAS_SIGNS, 16>(a, s);
}

}  // namespace

extern "C" int nemo_b200_skin_fwd(const float* A, const float* PO, const float* base, const float* Wt, float* V, int N,
                                  void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(A && PO && base && Wt && V && N >= 0, "skin_fwd: bad args");
  SkinArgs a{};
  a.A = A; a.PO = PO; a.base = base; a.Wt = Wt; a.V = V; a.N = N;
  chunking(N, FB, a.groups_per_chunk, a.nchunks);
  return launch_fwd<0>(a, S(stream));
}

extern "C" int nemo_b200_skin_l1_fwd(const float* A_a, const float* PO_a, const float* A_b, const float* PO_b,
                                     const float* base, const float* Wt, uint32_t* signs, float* l1_sum, int N,
                                     void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(A_a && PO_a && A_b && PO_b && base && Wt && signs && l1_sum && N >= 0, "skin_l1_fwd: bad args");
  SkinArgs a{};
  a.A = A_a; a.PO = PO_a; a.A2 = A_b; a.PO2 = PO_b; a.base = base; a.Wt = Wt; a.signs = signs; a.l1_sum = l1_sum;
  a.N = N;
  chunking(N, FB / 2, a.groups_per_chunk, a.nchunks);
  return launch_fwd<1>(a, S(stream));
}

extern "C" int nemo_b200_skin_bwd(const float* A, const float* PO, const float* base, const float* Wt, const float* dV,
                                  const uint32_t* signs, float sign_scale, float* dPO, float* dAt, int N,
                                  void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(A && PO && base && Wt && dPO && dAt && (dV || signs) && N >= 0, "skin_bwd: bad args");
  cudaStream_t s = S(stream);
  SkinBwdArgs a{};
  a.A = A; a.PO = PO; a.base = base; a.Wt = Wt; a.dV = dV; a.signs = signs;
  a.sign_scale = sign_scale; a.dPO = dPO; a.dAt = dAt; a.N = N;
  chunking(N, FB, a.groups_per_chunk, a.nchunks);
  static const int abl = [] { const char* e = getenv("NEMO_B200_SKIN_ABLATE"); return e ? atoi(e) : 0; }();
  a.ablate = abl;
  NB_CUDA(cudaMemsetAsync(dAt, 0, sizeof(float) * (size_t)N * 288, s));  // vertex tiles accumulate with reductions
  if (dV && signs) return launch_bwd<true, true>(a, s);
  if (dV) return launch_bwd<true, false>(a, s);
  return launch_bwd<false, true>(a, s);
}
