// K7 forward on the tensor cores: fused linear blend skinning + joint read-outs (+ fused vertex L1).
// Reference arithmetic: human_body_prior/body_model/lbs.py:236-252 (v_posed, T = W.A, v = T.[v_posed;1]),
// smplx VertexJointSelector (vertex picks), hmr/smpl.py:33-36 (J_regressor_extra),
// nemo/neural_motion_model.py:2792-2797 (vertex-to-vertex L1 of vposer_loss).
//
// With dense skinning weights the blend T[v] = sum_j W[v,j] A[j] costs 288 FMA per vertex-frame, 3x what the
// CUDA cores can do in the time HBM needs for the frame's streams.  Here it is a GEMM on tcgen05:
//   D[128 vertices, (frame, 12)] = W_tile[128, 24] . Amat[24, (frame, 12)]         (3xTF32: lo.hi + hi.lo + hi.hi)
// accumulated in TMEM, where lane = vertex: the thread that owns a vertex reads its 3x4 transform for every
// frame of the group straight from its TMEM lane, applies it to v_posed = base + pose_offsets (one coalesced
// planar read per component) and writes the vertex.  No per-vertex intermediate reaches HBM.
//
// Decomposition: a CTA owns one 128-vertex tile (W tile split once into tf32 hi/lo in shared memory, both
// operands in the canonical K-major 128-byte-swizzle layout) and walks a chunk of frame groups; four producer
// warps turn the group's A matrices into the B operand (double buffered) and issue the MMAs, four worker warps
// (one thread per vertex) consume T.  Cross-tile reductions (regressed joints, L1 sum) use fp32 atomics.
#include "common.cuh"
#include "sm100.cuh"

using namespace nb;
using namespace nb::sm100;

namespace {

constexpr int TILE_V = 128;
constexpr int NT_V = VP / TILE_V;  // 54
constexpr int WT_BYTES = TILE_V * 128;  // one tf32 part of the W tile: 128 rows x 32 k (24 used)
static_assert(VP % TILE_V == 0, "VP must be a multiple of the vertex tile");

// byte offset of element (row, k) in a K-major tile with 128-byte rows and the 128-byte swizzle
__device__ __forceinline__ uint32_t kmajor_sw128(int row, int k) {
  return (uint32_t)(row * 128 + ((((k >> 2) ^ row) & 7) << 4) + ((k & 3) << 2));
}

// MODE 0: forward, one mesh, groups of 8 frames.
// MODE 1: fused L1 of two meshes (mesh a | mesh b side by side in one accumulator), groups of 4 frames.
// Either way 96 accumulator columns, 81 KB of shared memory and 256 TMEM columns: 2 CTAs per SM.
template <int MODE>
struct SkinCfg {
  static constexpr int FB = MODE == 0 ? 8 : 4;
  static constexpr int NMESH = MODE == 0 ? 1 : 2;
  static constexpr int NCOL = NMESH * FB * 12;     // accumulator columns = B operand rows (96 / 192)
  static constexpr int BUF_COLS = NCOL <= 128 ? 128 : 256;  // TMEM column stride between the two accumulators
  static constexpr int AM_BYTES = NCOL * 128;      // one tf32 part of the B operand
  static constexpr int OFF_W = 0;                  // [W hi][W lo]
  static constexpr int OFF_AM = 2 * WT_BYTES;      // 2 buffers x [hi][lo]
  static constexpr int OFF_RED = OFF_AM + 4 * AM_BYTES;
  static constexpr int OFF_BARS = OFF_RED + 256;
  static constexpr int SMEM_BYTES = OFF_BARS + 64 + 1024;
  static constexpr int TMEM_COLS = 2 * BUF_COLS;
  static constexpr int CTAS_PER_SM = 2;
};

struct SkinArgs {
  const float* A;      // (N,24,12)   [MODE 1: mesh a]
  const float* PO;     // (N,3,VP) or NULL
  const float* A2;     // MODE 1: mesh b
  const float* PO2;
  const float* base;   // (3,VP)
  const float* Wt;     // (24,VP)
  const float* Jx;     // (nx,VP)
  const signed char* slot;
  float* V;            // (N,NV,3) or NULL
  float* JX;           // (N,np+nx,3)
  uint32_t* signs;     // MODE 1: (N,3,2,VP/32)
  float* l1_sum;
  int nx, np, N, groups_per_chunk, nchunks;
};

template <int MODE, int NXR>
__global__ void __launch_bounds__(256, SkinCfg<MODE>::CTAS_PER_SM) k_skin_tc_fwd(const SkinArgs a) {
  using C_ = SkinCfg<MODE>;
  constexpr int FB = C_::FB, NMESH = C_::NMESH, NCOL = C_::NCOL;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + C_::OFF_BARS;
  auto tfull = [&](int b) { return bars + 8u * b; };         // T of buffer b complete          (commit)
  auto tempty = [&](int b) { return bars + 8u * (2 + b); };  // workers done reading T[b]       (128)
  auto amfree = [&](int b) { return bars + 8u * (4 + b); };  // MMAs done reading Amat[b]       (commit)
  const uint32_t tmem_slot = bars + 48;
  const uint32_t sred = sbase + C_::OFF_RED;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int vt = blockIdx.x % NT_V, chunk = blockIdx.x / NT_V;
  const int v0 = vt * TILE_V;
  const int ngroups_total = (a.N + FB - 1) / FB;
  const int g_begin = chunk * a.groups_per_chunk;
  const int ng = min(ngroups_total, g_begin + a.groups_per_chunk) - g_begin;
  // vertex tiles walk the chunk's groups in rotated order so their atomics do not pile onto the same frames
  const int rot = (vt * 5) % ng;
  auto group_of = [&](int gi) { return g_begin + (gi + rot) % ng; };

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), 128);
      mbar_init(amfree(b), 1);
    }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<C_::TMEM_COLS>(tmem_slot);
  // zero both B-operand buffers once: the k = 24..31 pad columns stay zero for the whole kernel
  for (int i = tid; i < 4 * C_::AM_BYTES / 16; i += 256)
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
  constexpr uint32_t idesc = umma_idesc_tf32(128, NCOL, 0, 0);

  if (warp >= 4) {
    // ------------------------------------------------------------------ producers: Amat + MMA issue
    const int pt = tid - 128;
    for (int gi = 0; gi < ng; ++gi) {
      const int b = gi & 1;
      const uint32_t use = gi >> 1;  // how many times buffer b was used before
      if (use > 0) mbar_wait(amfree(b), (use - 1) & 1);
      const uint32_t am = sbase + C_::OFF_AM + b * 2 * C_::AM_BYTES;
      const int n0 = group_of(gi) * FB;
      // all loads of the group first (independent, L2 latency overlapped), then split + scatter into the operand
      constexpr int ITER = NMESH * FB * 288 / 128;
      float x[ITER];
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        const int e = pt + 128 * i;
        const int mesh = e / (FB * 288), r = e % (FB * 288), f = r / 288;
        const float* src = (NMESH == 2 && mesh == 1) ? a.A2 : a.A;
        x[i] = (n0 + f < a.N) ? __ldg(src + (size_t)n0 * 288 + r) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        const int e = pt + 128 * i;
        const int mesh = e / (FB * 288), r = e % (FB * 288);
        const int f = r / 288, jc = r % 288, j = jc / 12, c = jc % 12;
        float hi, lo;
        tf32_split(x[i], hi, lo);
        const uint32_t off = am + kmajor_sw128((mesh * FB + f) * 12 + c, j);
        sts32(off, hi);
        sts32(off + C_::AM_BYTES, lo);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (pt == 0) {
        if (use > 0) mbar_wait(tempty(b), (use - 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + b * C_::BUF_COLS;
        const uint64_t wh0 = umma_smem_desc(sbase + C_::OFF_W, 16, 1024, LAYOUT_SW128);
        const uint64_t mh0 = umma_smem_desc(sbase + C_::OFF_AM + b * 2 * C_::AM_BYTES, 16, 1024, LAYOUT_SW128);
#pragma unroll
        for (int j = 0; j < 3; ++j) {  // K = 24 = 3 steps of 8
          const uint64_t wh = umma_desc_advance(wh0, j * 32), wl = umma_desc_advance(wh0, WT_BYTES + j * 32);
          const uint64_t mh = umma_desc_advance(mh0, j * 32), ml = umma_desc_advance(mh0, C_::AM_BYTES + j * 32);
          umma_tf32(d, wl, mh, idesc, j > 0 ? 1u : 0u);
          umma_tf32(d, wh, ml, idesc, 1u);
          umma_tf32(d, wh, mh, idesc, 1u);
        }
        umma_commit(tfull(b));
        umma_commit(amfree(b));
      }
    }
  } else {
    // ------------------------------------------------------------------ workers: one thread per vertex
    const int v = v0 + tid;
    const bool real = v < NV;
    float bs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) bs[c] = __ldg(a.base + c * VP + v);
    float jx[NXR > 0 ? NXR : 1];
#pragma unroll
    for (int r = 0; r < NXR; ++r) jx[r] = r < a.nx ? __ldg(a.Jx + (size_t)r * VP + v) : 0.f;
    const int sl = a.slot ? a.slot[v] : -1;
    const int nxr = a.np + a.nx;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float l1 = 0.f;
    // pose offsets are prefetched one group ahead (registers), so HBM reads overlap the skinning arithmetic
    float po[NMESH][FB][3], pn[NMESH][FB][3];
    auto load_po = [&](int gi, float (&dst)[NMESH][FB][3]) {
      const int n0 = gi < ng ? group_of(gi) * FB : a.N;
#pragma unroll
      for (int m = 0; m < NMESH; ++m) {
        const float* P = (NMESH == 2 && m == 1) ? a.PO2 : a.PO;
#pragma unroll
        for (int f = 0; f < FB; ++f)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            dst[m][f][c] = (P && n0 + f < a.N) ? __ldcs(P + ((size_t)(n0 + f) * 3 + c) * VP + v) : 0.f;
      }
    };
    load_po(0, po);
    for (int gi = 0; gi < ng; ++gi) {
      const int b = gi & 1;
      const int n0 = group_of(gi) * FB;
      load_po(gi + 1, pn);
      mbar_wait(tfull(b), (gi >> 1) & 1);
      tc_fence_after();
      const uint32_t ta = tmem_base + lane_addr + b * C_::BUF_COLS;
#pragma unroll
      for (int f = 0; f < FB; ++f) {
        const int n = n0 + f;
        float o[NMESH][3];
#pragma unroll
        for (int m = 0; m < NMESH; ++m) {
          uint32_t t[12];
          tmem_ld4(ta + (m * FB + f) * 12, t);
          tmem_ld4(ta + (m * FB + f) * 12 + 4, t + 4);
          tmem_ld4(ta + (m * FB + f) * 12 + 8, t + 8);
          tmem_ld_wait();
          const float p0 = bs[0] + po[m][f][0], p1 = bs[1] + po[m][f][1], p2 = bs[2] + po[m][f][2];
#pragma unroll
          for (int r = 0; r < 3; ++r)
            o[m][r] = __uint_as_float(t[4 * r]) * p0 + __uint_as_float(t[4 * r + 1]) * p1 +
                      __uint_as_float(t[4 * r + 2]) * p2 + __uint_as_float(t[4 * r + 3]);
        }
        if (n < a.N) {  // warp-uniform
          if (MODE == 0) {
            if (a.V && real) {
              float* d = a.V + ((size_t)n * NV + v) * 3;
              __stcs(d, o[0][0]); __stcs(d + 1, o[0][1]); __stcs(d + 2, o[0][2]);
            }
            if (sl >= 0) {
              float* d = a.JX + ((size_t)n * nxr + sl) * 3;
              d[0] = o[0][0]; d[1] = o[0][1]; d[2] = o[0][2];
            }
            if (NXR > 0) {
              // regressed joints: per-warp partial sums over its 32 vertices, one lane per value, then one
              // vector of atomics per warp (the other vertex tiles add theirs)
              constexpr int NXP = NXR == 1 ? 4 : 32;
              float part[NXP];
#pragma unroll
              for (int k = 0; k < NXP; ++k) part[k] = 0.f;
#pragma unroll
              for (int r = 0; r < NXR; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) part[r * 3 + c] = jx[r] * o[0][c];
              warp_reduce_scatter<NXP>(part, lane);
              const int idx = scatter_index<NXP>(lane);
              const bool owner = NXP == 32 ? true : (lane & 7) == 0;
              if (owner && idx < 3 * a.nx) atomicAdd(a.JX + ((size_t)n * nxr + a.np) * 3 + idx, part[0]);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float dlt = real ? o[0][c] - o[NMESH - 1][c] : 0.f;
              l1 += fabsf(dlt);
              const unsigned pm = __ballot_sync(0xffffffffu, dlt > 0.f), nm = __ballot_sync(0xffffffffu, dlt < 0.f);
              if (lane == 0 && a.signs) {
                uint32_t* s = a.signs + ((size_t)n * 3 + c) * 2 * (VP / 32);
                s[v >> 5] = pm;
                s[(VP / 32) + (v >> 5)] = nm;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty(b));
#pragma unroll
      for (int m = 0; m < NMESH; ++m)
#pragma unroll
        for (int f = 0; f < FB; ++f)
#pragma unroll
          for (int c = 0; c < 3; ++c) po[m][f][c] = pn[m][f][c];
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
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<C_::TMEM_COLS>(tmem_base);
  }
}

// ================================================================================================ backward
// Per vertex tile and group of 8 frames, two GEMMs on tcgen05 (3xTF32):
//   T[v, (f,12)]    = W[v, :] . Amat[:, (f,12)]                       as in the forward (its 3x3 part gives dPO)
//   dA^T[(f,12), j] = sum_v G^T[(f,12), v] . W[v, j]                  G[v,(f,r,c)] = dv[f,v,r] * [v_posed;1][c]
// The worker thread of vertex v reads its T rows from TMEM, assembles dv (upstream dV + picked / regressed joint
// gradients + L1 sign bits), writes dPO = R^T dv (coalesced planar), and scatters the tf32 hi/lo parts of its
// 12 G values per frame into the K-major A operand of the second GEMM (row = (f,r,c), k = vertex: a warp writes
// one full 128-byte swizzled row per store, conflict free).  The dA^T accumulator (lanes = (f,r,c), columns =
// joints) is drained by three producer warps straight into global dA with fp32 atomics (54 vertex tiles add up).
struct BwdCfg {
  static constexpr int FB = 8;
  static constexpr int NROW = FB * 12;               // 96 rows: T columns / G^T rows
  static constexpr int AM_BYTES = NROW * 128;        // one tf32 part of Amat (K-major, 96 rows)
  static constexpr int WD_BYTES = 4 * 32 * 128;      // one tf32 part of W as dA-GEMM B operand: 4 k-blocks x 32 rows(j)
  static constexpr int GT_KB_BYTES = NROW * 128;     // one k-block (32 vertices) of one tf32 part of G^T
  static constexpr int GT_BYTES = 4 * GT_KB_BYTES;   // rows 96..127 of a block alias the next block (results unused)
  static constexpr int OFF_WT = 0;                               // [hi][lo] 2 x 16 KB
  static constexpr int OFF_WD = OFF_WT + 2 * WT_BYTES;           // [hi][lo] 2 x 16 KB
  static constexpr int OFF_AM = OFF_WD + 2 * WD_BYTES;           // 2 buffers x [hi][lo] x 12 KB
  static constexpr int OFF_GT = OFF_AM + 4 * AM_BYTES;           // [hi][lo] x 48 KB
  static constexpr int OFF_TAIL = OFF_GT + 2 * GT_BYTES;         // 4 KB of zeros read by the last block's rows 96..127
  static constexpr int OFF_DJ = OFF_TAIL + 4096;                 // joint gradients of the group: 2 x [8][96] floats
  static constexpr int OFF_BARS = OFF_DJ + 2 * FB * 96 * 4;
  static constexpr int SMEM_BYTES = OFF_BARS + 128 + 1024;
  static constexpr int TMEM_COLS = 512;              // T: 2 x 128, dA^T: 2 x 32 at column 256
};
static_assert(BwdCfg::SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct SkinBwdArgs {
  const float* A; const float* PO; const float* base; const float* Wt; const float* Jx;
  const signed char* slot;
  const float* dV; const float* dJX; const uint32_t* signs;
  float* dPO; float* dA;
  float sign_scale;
  int nx, np, N, groups_per_chunk, nchunks;
};

template <int NXR>
__global__ void __launch_bounds__(384, 1) k_skin_tc_bwd(const SkinBwdArgs a) {
  using C_ = BwdCfg;
  constexpr int FB = C_::FB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + C_::OFF_BARS;
  auto tfull = [&](int b) { return bars + 8u * b; };            // T[b] complete                 (commit)
  auto tempty = [&](int b) { return bars + 8u * (2 + b); };     // workers done with T[b]        (128)
  auto amfree = [&](int b) { return bars + 8u * (4 + b); };     // T-GEMM done reading Amat[b]   (commit)
  auto d2full = [&](int b) { return bars + 8u * (6 + b); };     // dA^T[b] complete              (commit)
  auto d2empty = [&](int b) { return bars + 8u * (8 + b); };    // dA^T[b] drained               (96)
  const uint32_t gfull = bars + 80;                             // G^T written                   (128)
  const uint32_t gfree = bars + 88;                             // dA-GEMM done reading G^T      (commit)
  auto djfull = [&](int b) { return bars + 8u * (12 + b); };    // joint gradients staged        (96)
  const uint32_t tmem_slot = bars + 112;
  const uint32_t sdj = sbase + C_::OFF_DJ;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int vt = blockIdx.x % NT_V, chunk = blockIdx.x / NT_V;
  const int v0 = vt * TILE_V;
  const int ngroups_total = (a.N + FB - 1) / FB;
  const int g_begin = chunk * a.groups_per_chunk;
  const int ng = min(ngroups_total, g_begin + a.groups_per_chunk) - g_begin;
  // vertex tiles walk the chunk's groups in rotated order so their dA atomics do not pile onto the same frames
  const int rot = (vt * 5) % ng;
  auto group_of = [&](int gi) { return g_begin + (gi + rot) % ng; };

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), 256);
      mbar_init(amfree(b), 1);
      mbar_init(d2full(b), 1);
      mbar_init(d2empty(b), 96);
      mbar_init(djfull(b), 96);
    }
    mbar_init(gfull, 256);
    mbar_init(gfree, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<C_::TMEM_COLS>(tmem_slot);
  // zero Amat (pad columns k = 24..31), G^T and its tail
  for (int i = tid; i < (C_::OFF_DJ - C_::OFF_AM) / 16; i += 384)
    sts128(sbase + C_::OFF_AM + 16 * i, make_float4(0.f, 0.f, 0.f, 0.f));
  if (tid < TILE_V) {
    const int v = v0 + tid;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      float w = k < NJ ? __ldg(a.Wt + (size_t)k * VP + v) : 0.f, hi, lo;
      tf32_split(w, hi, lo);
      // T-GEMM A operand: row = vertex, k = joint
      const uint32_t o1 = sbase + C_::OFF_WT + kmajor_sw128(tid, k);
      sts32(o1, hi);
      sts32(o1 + WT_BYTES, lo);
      // dA-GEMM B operand: row = joint, k = vertex (k-block = warp, 32 rows x 128 bytes each)
      const uint32_t o2 = sbase + C_::OFF_WD + (uint32_t)(warp * 4096) + kmajor_sw128(k, lane);
      sts32(o2, hi);
      sts32(o2 + C_::WD_BYTES, lo);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32u(tmem_slot);
  constexpr uint32_t idesc_t = umma_idesc_tf32(128, C_::NROW, 0, 0);
  constexpr uint32_t idesc_d = umma_idesc_tf32(128, 32, 0, 0);

  if (warp >= 4 && warp < 7) {
    // ------------------------------------------------------------------ producers: Amat, T-GEMM issue, dA^T drain
    const int pt = tid - 128;
    auto drain = [&](int gi) {
      const int b = gi & 1;
      mbar_wait(d2full(b), (gi >> 1) & 1);
      tc_fence_after();
      uint32_t t[24];
      const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256 + b * 32;
      tmem_ld16(ta, t);
      tmem_ld4(ta + 16, t + 16);
      tmem_ld4(ta + 20, t + 20);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(d2empty(b));
      const int row = (warp & 3) * 32 + lane;  // (f, r, c)
      const int n = group_of(gi) * FB + row / 12;
      if (n < a.N) {
        float* d = a.dA + (size_t)n * 288 + row % 12;
#pragma unroll
        for (int j = 0; j < NJ; ++j) atomicAdd(d + j * 12, __uint_as_float(t[j]));
      }
    };
    for (int gi = 0; gi < ng; ++gi) {
      const int b = gi & 1;
      const uint32_t use = gi >> 1;
      if (use > 0) mbar_wait(amfree(b), (use - 1) & 1);
      const uint32_t am = sbase + C_::OFF_AM + b * 2 * C_::AM_BYTES;
      const int n0 = group_of(gi) * FB;
      constexpr int ITER = FB * 288 / 96;
      float x[ITER], dj[FB];
      const int nxr3 = (a.np + a.nx) * 3;
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        const int e = pt + 96 * i;
        x[i] = (n0 + e / 288 < a.N) ? __ldg(a.A + (size_t)n0 * 288 + e) : 0.f;
      }
#pragma unroll
      for (int f = 0; f < FB; ++f)
        dj[f] = (a.dJX && pt < nxr3 && n0 + f < a.N) ? __ldg(a.dJX + (size_t)(n0 + f) * nxr3 + pt) : 0.f;
#pragma unroll
      for (int f = 0; f < FB; ++f) sts32(sdj + 4 * ((b * FB + f) * 96 + pt), dj[f]);
      mbar_arrive(djfull(b));
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        const int e = pt + 96 * i;
        const int f = e / 288, jc = e % 288, j = jc / 12, c = jc % 12;
        float hi, lo;
        tf32_split(x[i], hi, lo);
        const uint32_t off = am + kmajor_sw128(f * 12 + c, j);
        sts32(off, hi);
        sts32(off + C_::AM_BYTES, lo);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 96;" ::: "memory");
      if (pt == 0) {
        if (use > 0) mbar_wait(tempty(b), (use - 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + b * 128;
        const uint64_t wh0 = umma_smem_desc(sbase + C_::OFF_WT, 16, 1024, LAYOUT_SW128);
        const uint64_t mh0 = umma_smem_desc(sbase + C_::OFF_AM + b * 2 * C_::AM_BYTES, 16, 1024, LAYOUT_SW128);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const uint64_t wh = umma_desc_advance(wh0, j * 32), wl = umma_desc_advance(wh0, WT_BYTES + j * 32);
          const uint64_t mh = umma_desc_advance(mh0, j * 32), ml = umma_desc_advance(mh0, C_::AM_BYTES + j * 32);
          umma_tf32(d, wl, mh, idesc_t, j > 0 ? 1u : 0u);
          umma_tf32(d, wh, ml, idesc_t, 1u);
          umma_tf32(d, wh, mh, idesc_t, 1u);
        }
        umma_commit(tfull(b));
        umma_commit(amfree(b));
      }
      if (gi > 0) drain(gi - 1);
    }
    drain(ng - 1);
  } else if (warp == 7) {
    // ------------------------------------------------------------------ dA-GEMM issuer
    if (lane == 0) {
      const uint64_t gh0 = umma_smem_desc(sbase + C_::OFF_GT, 16, 1024, LAYOUT_SW128);
      const uint64_t wd0 = umma_smem_desc(sbase + C_::OFF_WD, 16, 1024, LAYOUT_SW128);
      for (int gi = 0; gi < ng; ++gi) {
        const int b = gi & 1;
        const uint32_t use = gi >> 1;
        mbar_wait(gfull, gi & 1);
        if (use > 0) mbar_wait(d2empty(b), (use - 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + 256 + b * 32;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t go = kb * C_::GT_KB_BYTES + ks * 32, wo = kb * 4096 + ks * 32;
            const uint64_t gh = umma_desc_advance(gh0, go), gl = umma_desc_advance(gh0, C_::GT_BYTES + go);
            const uint64_t wh = umma_desc_advance(wd0, wo), wl = umma_desc_advance(wd0, C_::WD_BYTES + wo);
            umma_tf32(d, gl, wh, idesc_d, (kb | ks) ? 1u : 0u);
            umma_tf32(d, gh, wl, idesc_d, 1u);
            umma_tf32(d, gh, wh, idesc_d, 1u);
          }
        }
        umma_commit(d2full(b));
        umma_commit(gfree);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ workers: two threads per vertex
    // warps 0-3 take frames 0-3 of every group, warps 8-11 frames 4-7 (same TMEM lane quarter = warp & 3):
    // two worker warps per scheduler, so one hides the other's TMEM / shared-memory latencies
    constexpr int FH = FB / 2;
    const int wq = warp & 3, half = warp >> 3;
    const int vl = wq * 32 + lane;  // vertex within the tile
    const int v = v0 + vl;
    const bool real = v < NV;
    float bs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) bs[c] = __ldg(a.base + c * VP + v);
    float jx[NXR > 0 ? NXR : 1];
#pragma unroll
    for (int r = 0; r < NXR; ++r) jx[r] = r < a.nx ? __ldg(a.Jx + (size_t)r * VP + v) : 0.f;
    const int sl = a.slot ? a.slot[v] : -1;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const uint32_t gt = sbase + C_::OFF_GT + wq * C_::GT_KB_BYTES;  // this quarter's k-block of G^T (hi part)
    float po[FH][3], pn[FH][3], dv[FH][3], dn[FH][3];
    uint32_t sw = 0u, swn = 0u;  // this half's 24 sign words, one per lane
    auto load_in = [&](int gi, float (&p)[FH][3], float (&d)[FH][3], uint32_t& w) {
      const int n0 = gi < ng ? group_of(gi) * FB + half * FH : a.N;
      if (a.signs) {
        const int fc = lane >> 1;  // word (f*3 + c)*2 + {pos, neg}
        const int n = n0 + fc / 3;
        w = (lane < FH * 6 && n < a.N)
                ? __ldg(a.signs + (((size_t)n * 3 + fc % 3) * 2 + (lane & 1)) * (VP / 32) + (v >> 5)) : 0u;
      }
#pragma unroll
      for (int f = 0; f < FH; ++f) {
        const int n = n0 + f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          p[f][c] = (a.PO && n < a.N) ? __ldcs(a.PO + ((size_t)n * 3 + c) * VP + v) : 0.f;
          d[f][c] = (a.dV && real && n < a.N) ? __ldcs(a.dV + ((size_t)n * NV + v) * 3 + c) : 0.f;
        }
      }
    };
    load_in(0, po, dv, sw);
    for (int gi = 0; gi < ng; ++gi) {
      const int b = gi & 1;
      const int n0 = group_of(gi) * FB + half * FH;
      load_in(gi + 1, pn, dn, swn);
      if (a.dJX) mbar_wait(djfull(b), (gi >> 1) & 1);
      // ---- upstream gradient of every vertex of the group: dV + sign bits + joint read-outs
#pragma unroll
      for (int f = 0; f < FH; ++f) {
        const int n = n0 + f;
        if (n < a.N) {  // warp-uniform
          if (a.signs) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const uint32_t pm = __shfl_sync(0xffffffffu, sw, (f * 3 + c) * 2);
              const uint32_t nm = __shfl_sync(0xffffffffu, sw, (f * 3 + c) * 2 + 1);
              dv[f][c] += a.sign_scale * ((float)((pm >> lane) & 1u) - (float)((nm >> lane) & 1u));
            }
          }
          if (a.dJX) {
            const uint32_t dj = sdj + 4 * ((b * FB + half * FH + f) * 96);
            if (sl >= 0) { dv[f][0] += lds32(dj + 12 * sl); dv[f][1] += lds32(dj + 12 * sl + 4); dv[f][2] += lds32(dj + 12 * sl + 8); }
#pragma unroll
            for (int r = 0; r < NXR; ++r) {
              if (r < a.nx) {
#pragma unroll
                for (int c = 0; c < 3; ++c) dv[f][c] += jx[r] * lds32(dj + 4 * ((a.np + r) * 3 + c));
              }
            }
          }
        }
      }
      // ---- dPO = R^T dv from the T rows in TMEM (all of this half's rows in one batch of loads)
      mbar_wait(tfull(b), (gi >> 1) & 1);
      tc_fence_after();
      const uint32_t ta = tmem_base + lane_addr + b * 128 + half * (FH * 12);
      uint32_t t[FH * 12];
      tmem_ld16(ta, t);
      tmem_ld16(ta + 16, t + 16);
      tmem_ld16(ta + 32, t + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(tempty(b));
#pragma unroll
      for (int f = 0; f < FH; ++f) {
        const int n = n0 + f;
        if (a.dPO && n < a.N) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            __stcs(a.dPO + ((size_t)n * 3 + c) * VP + v,
                   __uint_as_float(t[f * 12 + c]) * dv[f][0] + __uint_as_float(t[f * 12 + 4 + c]) * dv[f][1] +
                       __uint_as_float(t[f * 12 + 8 + c]) * dv[f][2]);
        }
      }
      // ---- G^T for the dA GEMM (the previous group's GEMM must have finished reading it)
      if (gi > 0) mbar_wait(gfree, (gi - 1) & 1);
#pragma unroll
      for (int f = 0; f < FH; ++f) {
        const float ph[4] = {bs[0] + po[f][0], bs[1] + po[f][1], bs[2] + po[f][2], 1.f};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float hi, lo;
            tf32_split(dv[f][r] * ph[c], hi, lo);
            const uint32_t off = gt + kmajor_sw128((half * FH + f) * 12 + r * 4 + c, lane);
            sts32(off, hi);
            sts32(off + C_::GT_BYTES, lo);
          }
      }
      fence_proxy_async_smem();
      mbar_arrive(gfull);
#pragma unroll
      for (int f = 0; f < FH; ++f)
#pragma unroll
        for (int c = 0; c < 3; ++c) { po[f][c] = pn[f][c]; dv[f][c] = dn[f][c]; }
      sw = swn;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
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
template <int FB>
void chunking(int N, int& groups_per_chunk, int& nchunks) {
  const int groups = cdiv(N, FB);
  int want = max(1, (8 * num_sms()) / NT_V);  // chunks so that 54 * chunks ~ 8 waves
  groups_per_chunk = max(min(4, groups), cdiv(groups, want));
  nchunks = cdiv(groups, groups_per_chunk);
}

template <int MODE, int NXR>
int launch(const SkinArgs& a, cudaStream_t s) {
  using C_ = SkinCfg<MODE>;
  static bool configured = false;
  if (!configured) {
    NB_CUDA(cudaFuncSetAttribute(k_skin_tc_fwd<MODE, NXR>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES));
    configured = true;
  }
  k_skin_tc_fwd<MODE, NXR><<<NT_V * a.nchunks, 256, C_::SMEM_BYTES, s>>>(a);
  NB_LAUNCH_CHECK("skin_tc_fwd");
}

}  // namespace

extern "C" int nemo_b200_skin_fwd(const float* A, const float* PO, const float* base, const float* Wt, const float* Jx,
                                  int nx, const signed char* slot, int np, float* V, float* JX, int N, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(A && base && Wt && N >= 0 && nx >= 0 && nx <= 10 && np >= 0 && np <= 22, "skin_fwd: bad args");
  NB_REQUIRE((nx == 0 || Jx) && (np == 0 || slot) && (nx + np == 0 || JX), "skin_fwd: joint read-out pointers missing");
  cudaStream_t s = S(stream);
  SkinArgs a{};
  a.A = A; a.PO = PO; a.base = base; a.Wt = Wt; a.Jx = Jx; a.slot = np ? slot : nullptr; a.V = V; a.JX = JX;
  a.nx = nx; a.np = np; a.N = N;
  chunking<SkinCfg<0>::FB>(N, a.groups_per_chunk, a.nchunks);
  // regressed joints are accumulated across vertex tiles with atomics
  if (nx > 0) NB_CUDA(cudaMemsetAsync(JX, 0, sizeof(float) * (size_t)N * (np + nx) * 3, s));
  if (nx == 0) return launch<0, 0>(a, s);
  if (nx == 1) return launch<0, 1>(a, s);
  return launch<0, 10>(a, s);
}

extern "C" int nemo_b200_skin_l1_fwd(const float* A_a, const float* PO_a, const float* A_b, const float* PO_b,
                                     const float* base, const float* Wt, uint32_t* signs, float* l1_sum, int N,
                                     void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(A_a && PO_a && A_b && PO_b && base && Wt && l1_sum && N >= 0, "skin_l1_fwd: bad args");
  SkinArgs a{};
  a.A = A_a; a.PO = PO_a; a.A2 = A_b; a.PO2 = PO_b; a.base = base; a.Wt = Wt; a.signs = signs; a.l1_sum = l1_sum;
  a.N = N;
  chunking<SkinCfg<1>::FB>(N, a.groups_per_chunk, a.nchunks);
  return launch<1, 0>(a, S(stream));
}

template <int NXR>
static int launch_bwd(const SkinBwdArgs& a, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    NB_CUDA(cudaFuncSetAttribute(k_skin_tc_bwd<NXR>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdCfg::SMEM_BYTES));
    configured = true;
  }
  k_skin_tc_bwd<NXR><<<NT_V * a.nchunks, 384, BwdCfg::SMEM_BYTES, s>>>(a);
  NB_LAUNCH_CHECK("skin_tc_bwd");
}

extern "C" int nemo_b200_skin_bwd(const float* A, const float* PO, const float* base, const float* Wt, const float* W2,
                                  const float* Jx, int nx, const signed char* slot, int np, const float* dV,
                                  const float* dJX, const uint32_t* signs, float sign_scale, float* dPO, float* dA,
                                  int N, void* stream) {
  (void)W2;
  if (N == 0) return 0;
  NB_REQUIRE(A && base && Wt && dA && N >= 0 && nx >= 0 && nx <= 10 && np >= 0 && np <= 22, "skin_bwd: bad args");
  NB_REQUIRE(!dJX || ((nx == 0 || Jx) && (np == 0 || slot)), "skin_bwd: joint read-out pointers missing");
  cudaStream_t s = S(stream);
  SkinBwdArgs a{};
  a.A = A; a.PO = PO; a.base = base; a.Wt = Wt; a.Jx = Jx; a.dV = dV; a.dJX = dJX; a.signs = signs;
  a.sign_scale = sign_scale; a.dPO = dPO; a.dA = dA; a.N = N;
  a.nx = dJX ? nx : 0;
  a.np = np;
  a.slot = (dJX && np) ? slot : nullptr;
  chunking<BwdCfg::FB>(N, a.groups_per_chunk, a.nchunks);
  NB_CUDA(cudaMemsetAsync(dA, 0, sizeof(float) * (size_t)N * 288, s));  // vertex tiles accumulate with atomics
  if (a.nx == 0) return launch_bwd<0>(a, s);
  if (a.nx == 1) return launch_bwd<1>(a, s);
  return launch_bwd<10>(a, s);
}
