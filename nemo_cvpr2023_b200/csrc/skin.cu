// K7: fused linear blend skinning + joint read-outs (+ fused vertex L1), forward and backward.
// Reference arithmetic: human_body_prior/body_model/lbs.py:236-252 (v_posed, T = W.A, v = T.[v_posed;1]),
// smplx VertexJointSelector (vertex picks), hmr/smpl.py:33-36 (J_regressor_extra),
// nemo/neural_motion_model.py:2792-2797 (vertex-to-vertex L1 of vposer_loss).
//
// Orientation: a CTA OWNS a group of FB frames and sweeps all 27 vertex tiles (256 vertices, two per
// thread), so every cross-vertex reduction (regressed joints, dA) stays inside the CTA: no global
// atomics, deterministic, and no per-vertex intermediate (T, v_posed, v_homo ...) ever reaches HBM.
// Per-frame inputs are read once from HBM in the planar [3][VP] layout (fully coalesced); the skinning
// weights (0.66 MB) are re-read per frame group from L2.
#include "common.cuh"

using namespace nb;

namespace {

constexpr int SK_THREADS = 128;
constexpr int SK_TILE = 256;
constexpr int SK_NT = VP / SK_TILE;  // 27
constexpr int SG_STRIDE = 20;        // floats per vertex row of the G staging tile (16 used; 20 => conflict-free)
static_assert(VP % SK_TILE == 0, "VP must be a multiple of the tile");

// Reduce K (power of two <= 32) per-lane values across the warp; on return v[0] of lane l holds the
// total of value index (l & (K-1)) ... for K == 32 that is value l.  31 shuffles for K = 32.
template <int K>
__device__ __forceinline__ void warp_reduce_scatter(float* v, int lane) {
  int off = 16;
#pragma unroll
  for (int half = K / 2; half >= 1; half >>= 1, off >>= 1) {
    bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      float send = upper ? v[i] : v[i + half];
      float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}
// index of the value that lane ends up holding after warp_reduce_scatter<K>
template <int K>
__device__ __forceinline__ int scatter_index(int lane) {
  if (K == 32) return lane;
  if (K == 4) return ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
  return 0;
}

__device__ __forceinline__ void accum_T(const float* __restrict__ sAf, const float* w, float* T) {
#pragma unroll
  for (int k = 0; k < 12; ++k) T[k] = 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const float4* a = reinterpret_cast<const float4*>(sAf + 12 * j);
    float4 a0 = a[0], a1 = a[1], a2 = a[2];
    float wj = w[j];
    T[0] += wj * a0.x; T[1] += wj * a0.y; T[2] += wj * a0.z; T[3] += wj * a0.w;
    T[4] += wj * a1.x; T[5] += wj * a1.y; T[6] += wj * a1.z; T[7] += wj * a1.w;
    T[8] += wj * a2.x; T[9] += wj * a2.y; T[10] += wj * a2.z; T[11] += wj * a2.w;
  }
}
// two vertices at once: halves the shared-memory traffic per FMA
__device__ __forceinline__ void accum_T2(const float* __restrict__ sAf, const float* w0, const float* w1, float* T0,
                                         float* T1) {
#pragma unroll
  for (int k = 0; k < 12; ++k) { T0[k] = 0.f; T1[k] = 0.f; }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const float4* a = reinterpret_cast<const float4*>(sAf + 12 * j);
    float4 a0 = a[0], a1 = a[1], a2 = a[2];
    float u = w0[j], v = w1[j];
    T0[0] += u * a0.x; T0[1] += u * a0.y; T0[2] += u * a0.z; T0[3] += u * a0.w;
    T0[4] += u * a1.x; T0[5] += u * a1.y; T0[6] += u * a1.z; T0[7] += u * a1.w;
    T0[8] += u * a2.x; T0[9] += u * a2.y; T0[10] += u * a2.z; T0[11] += u * a2.w;
    T1[0] += v * a0.x; T1[1] += v * a0.y; T1[2] += v * a0.z; T1[3] += v * a0.w;
    T1[4] += v * a1.x; T1[5] += v * a1.y; T1[6] += v * a1.z; T1[7] += v * a1.w;
    T1[8] += v * a2.x; T1[9] += v * a2.y; T1[10] += v * a2.z; T1[11] += v * a2.w;
  }
}
// rotation part only (backward needs T.R^T dv, not the translation)
__device__ __forceinline__ void accum_R2(const float* __restrict__ sAf, const float* w0, const float* w1, float* T0,
                                         float* T1) {
#pragma unroll
  for (int k = 0; k < 9; ++k) { T0[k] = 0.f; T1[k] = 0.f; }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const float4* a = reinterpret_cast<const float4*>(sAf + 12 * j);
    float4 a0 = a[0], a1 = a[1], a2 = a[2];
    float u = w0[j], v = w1[j];
    T0[0] += u * a0.x; T0[1] += u * a0.y; T0[2] += u * a0.z;
    T0[3] += u * a1.x; T0[4] += u * a1.y; T0[5] += u * a1.z;
    T0[6] += u * a2.x; T0[7] += u * a2.y; T0[8] += u * a2.z;
    T1[0] += v * a0.x; T1[1] += v * a0.y; T1[2] += v * a0.z;
    T1[3] += v * a1.x; T1[4] += v * a1.y; T1[5] += v * a1.z;
    T1[6] += v * a2.x; T1[7] += v * a2.y; T1[8] += v * a2.z;
  }
}
__device__ __forceinline__ void apply_T(const float* T, const float* p, float* o) {
#pragma unroll
  for (int r = 0; r < 3; ++r) o[r] = T[4 * r] * p[0] + T[4 * r + 1] * p[1] + T[4 * r + 2] * p[2] + T[4 * r + 3];
}

__device__ __forceinline__ void load_frames(float* sA, const float* __restrict__ A, int n0, int nf, int N) {
  for (int i = threadIdx.x; i < nf * 72; i += SK_THREADS) {
    int f = i / 72;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + f < N) v = reinterpret_cast<const float4*>(A + (size_t)(n0 + f) * 288)[i % 72];
    reinterpret_cast<float4*>(sA)[i] = v;
  }
}

// ------------------------------------------------------------------------------------ forward
// NXP: padded count of regressed-joint partial sums per warp reduction (0, 4 -> 1 row, 32 -> up to 10 rows)
template <int FB, int NXP, bool WRITE_V>
__global__ void __launch_bounds__(SK_THREADS) k_skin_fwd(const float* __restrict__ A, const float* __restrict__ PO,
                                                         const float* __restrict__ base, const float* __restrict__ Wt,
                                                         const float* __restrict__ Jx, int nx,
                                                         const signed char* __restrict__ slot, int np,
                                                         float* __restrict__ V, float* __restrict__ JX, int N) {
  constexpr int NXR = NXP == 32 ? 10 : (NXP == 4 ? 1 : 0);
  __shared__ __align__(16) float sA[FB * 288];
  __shared__ float sJ[FB * 32];
  const int tid = threadIdx.x, lane = tid & 31;
  const int n0 = blockIdx.x * FB;
  const int nxr = np + nx;
  load_frames(sA, A, n0, FB, N);
  for (int i = tid; i < FB * 32; i += SK_THREADS) sJ[i] = 0.f;
  __syncthreads();
  for (int t = 0; t < SK_NT; ++t) {
    const int v0 = t * SK_TILE + tid, v1 = v0 + SK_THREADS;
    float w0[NJ], w1[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) { w0[j] = __ldg(Wt + (size_t)j * VP + v0); w1[j] = __ldg(Wt + (size_t)j * VP + v1); }
    float b0[3], b1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { b0[c] = __ldg(base + c * VP + v0); b1[c] = __ldg(base + c * VP + v1); }
    float jx0[NXR > 0 ? NXR : 1], jx1[NXR > 0 ? NXR : 1];
#pragma unroll
    for (int r = 0; r < NXR; ++r) {
      jx0[r] = r < nx ? __ldg(Jx + (size_t)r * VP + v0) : 0.f;
      jx1[r] = r < nx ? __ldg(Jx + (size_t)r * VP + v1) : 0.f;
    }
    const int s0 = slot ? slot[v0] : -1, s1 = slot ? slot[v1] : -1;
#pragma unroll
    for (int f = 0; f < FB; ++f) {
      const int n = n0 + f;
      if (n >= N) break;
      float T0[12], T1[12];
      accum_T2(sA + f * 288, w0, w1, T0, T1);
      float p0[3] = {b0[0], b0[1], b0[2]}, p1[3] = {b1[0], b1[1], b1[2]};
      if (PO) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          p0[c] += __ldcs(PO + ((size_t)n * 3 + c) * VP + v0);
          p1[c] += __ldcs(PO + ((size_t)n * 3 + c) * VP + v1);
        }
      }
      float o0[3], o1[3];
      apply_T(T0, p0, o0);
      apply_T(T1, p1, o1);
      if (WRITE_V) {
        if (v0 < NV) {
          float* d = V + ((size_t)n * NV + v0) * 3;
          __stcs(d, o0[0]); __stcs(d + 1, o0[1]); __stcs(d + 2, o0[2]);
        }
        if (v1 < NV) {
          float* d = V + ((size_t)n * NV + v1) * 3;
          __stcs(d, o1[0]); __stcs(d + 1, o1[1]); __stcs(d + 2, o1[2]);
        }
      }
      if (s0 >= 0) { float* d = JX + ((size_t)n * nxr + s0) * 3; d[0] = o0[0]; d[1] = o0[1]; d[2] = o0[2]; }
      if (s1 >= 0) { float* d = JX + ((size_t)n * nxr + s1) * 3; d[0] = o1[0]; d[1] = o1[1]; d[2] = o1[2]; }
      if (NXP > 0) {
        float part[NXP > 0 ? NXP : 1];
#pragma unroll
        for (int k = 0; k < NXP; ++k) part[k] = 0.f;
#pragma unroll
        for (int r = 0; r < NXR; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) part[r * 3 + c] = jx0[r] * o0[c] + jx1[r] * o1[c];
        warp_reduce_scatter<NXP>(part, lane);
        int idx = scatter_index<NXP>(lane);
        bool owner = NXP == 32 ? true : (lane & 7) == 0;
        if (owner && idx < 3 * nx) atomicAdd(&sJ[f * 32 + idx], part[0]);
      }
    }
  }
  if (NXP > 0) {
    __syncthreads();
    for (int i = tid; i < FB * 3 * nx; i += SK_THREADS) {
      int f = i / (3 * nx), k = i % (3 * nx);
      if (n0 + f < N) JX[((size_t)(n0 + f) * nxr + np) * 3 + k] = sJ[f * 32 + k];
    }
  }
}

// ------------------------------------------------------------------------------------ fused L1 forward
// sum |v_b - v_a| over real vertices; sign bits of (v_a - v_b) are kept (6 bits per vertex-frame) so the
// backward never touches mesh b again.  signs: [N][3][2][VP/32] (pos, neg masks).
template <int FB>
__global__ void __launch_bounds__(SK_THREADS) k_skin_l1_fwd(const float* __restrict__ Aa, const float* __restrict__ POa,
                                                            const float* __restrict__ Ab, const float* __restrict__ POb,
                                                            const float* __restrict__ base, const float* __restrict__ Wt,
                                                            uint32_t* __restrict__ signs, float* __restrict__ l1_sum,
                                                            int N) {
  __shared__ __align__(16) float sA[2 * FB * 288];
  __shared__ float sred[SK_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * FB;
  load_frames(sA, Aa, n0, FB, N);
  load_frames(sA + FB * 288, Ab, n0, FB, N);
  __syncthreads();
  float total = 0.f;
  for (int t = 0; t < SK_NT; ++t) {
    const int v0 = t * SK_TILE + tid, v1 = v0 + SK_THREADS;
    float w0[NJ], w1[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) { w0[j] = __ldg(Wt + (size_t)j * VP + v0); w1[j] = __ldg(Wt + (size_t)j * VP + v1); }
    float b0[3], b1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { b0[c] = __ldg(base + c * VP + v0); b1[c] = __ldg(base + c * VP + v1); }
#pragma unroll
    for (int f = 0; f < FB; ++f) {
      const int n = n0 + f;
      if (n >= N) break;
      float oa0[3], oa1[3], ob0[3], ob1[3];
      {
        float T0[12], T1[12], p0[3], p1[3];
        accum_T2(sA + f * 288, w0, w1, T0, T1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          p0[c] = b0[c] + __ldcs(POa + ((size_t)n * 3 + c) * VP + v0);
          p1[c] = b1[c] + __ldcs(POa + ((size_t)n * 3 + c) * VP + v1);
        }
        apply_T(T0, p0, oa0);
        apply_T(T1, p1, oa1);
      }
      {
        float T0[12], T1[12], p0[3], p1[3];
        accum_T2(sA + (FB + f) * 288, w0, w1, T0, T1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          p0[c] = b0[c] + __ldcs(POb + ((size_t)n * 3 + c) * VP + v0);
          p1[c] = b1[c] + __ldcs(POb + ((size_t)n * 3 + c) * VP + v1);
        }
        apply_T(T0, p0, ob0);
        apply_T(T1, p1, ob1);
      }
      const bool r0 = v0 < NV, r1 = v1 < NV;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float d0 = r0 ? oa0[c] - ob0[c] : 0.f, d1 = r1 ? oa1[c] - ob1[c] : 0.f;
        total += fabsf(d0) + fabsf(d1);
        unsigned p0m = __ballot_sync(0xffffffffu, d0 > 0.f), n0m = __ballot_sync(0xffffffffu, d0 < 0.f);
        unsigned p1m = __ballot_sync(0xffffffffu, d1 > 0.f), n1m = __ballot_sync(0xffffffffu, d1 < 0.f);
        if (lane == 0 && signs) {
          uint32_t* s = signs + ((size_t)n * 3 + c) * 2 * (VP / 32);
          int wd0 = t * 8 + warp, wd1 = wd0 + 4;
          s[wd0] = p0m; s[(VP / 32) + wd0] = n0m;
          s[wd1] = p1m; s[(VP / 32) + wd1] = n1m;
        }
      }
    }
  }
  total = warp_sum(total);
  if (lane == 0) sred[warp] = total;
  __syncthreads();
  if (tid == 0) atomicAdd(l1_sum, sred[0] + sred[1] + sred[2] + sred[3]);
}

// ------------------------------------------------------------------------------------ backward
// Part 1 (thread = 2 vertices): dv assembly, dPO = T.R^T dv, G = dv (x) [v_posed;1] staged in shared memory.
// Part 2 (lane = 3 joints x 3 components, warp = 64-vertex slice): dA[j][c] += W[v][j] G[v][c] accumulated in
// registers across all 27 tiles, so dA needs no atomics.
template <int FB, int NXR>
__global__ void __launch_bounds__(SK_THREADS) k_skin_bwd(const float* __restrict__ A, const float* __restrict__ PO,
                                                         const float* __restrict__ base, const float* __restrict__ Wt,
                                                         const float* __restrict__ W2, const float* __restrict__ Jx,
                                                         int nx, const signed char* __restrict__ slot, int np,
                                                         const float* __restrict__ dV, const float* __restrict__ dJX,
                                                         const uint32_t* __restrict__ signs, float sign_scale,
                                                         float* __restrict__ dPO, float* __restrict__ dA, int N) {
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;                               // FB*288
  float* sdJ = sA + FB * 288;                     // FB*96
  float* sW2 = sdJ + FB * 96;                     // 256*32
  float* sG = sW2 + SK_TILE * 32;                 // 2*256*SG_STRIDE
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int jt = lane >> 2, cq = lane & 3;
  const int n0 = blockIdx.x * FB;
  const int nxr = np + nx;
  load_frames(sA, A, n0, FB, N);
  for (int i = tid; i < FB * 96; i += SK_THREADS) {
    int f = i / 96, k = i % 96;
    sdJ[i] = (dJX && n0 + f < N && k < nxr * 3) ? dJX[(size_t)(n0 + f) * nxr * 3 + k] : 0.f;
  }
  float acc[FB][9];
#pragma unroll
  for (int f = 0; f < FB; ++f)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[f][k] = 0.f;

  for (int t = 0; t < SK_NT; ++t) {
    __syncthreads();  // previous tile's part 2 is done with sW2 / sG (also publishes sA, sdJ on t == 0)
    {
      const float4* src = reinterpret_cast<const float4*>(W2 + (size_t)t * SK_TILE * 32);
      float4* dst = reinterpret_cast<float4*>(sW2);
#pragma unroll
      for (int i = 0; i < (SK_TILE * 32 / 4) / SK_THREADS; ++i) dst[tid + i * SK_THREADS] = __ldg(src + tid + i * SK_THREADS);
    }
    const int v0 = t * SK_TILE + tid, v1 = v0 + SK_THREADS;
    float w0[NJ], w1[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) { w0[j] = __ldg(Wt + (size_t)j * VP + v0); w1[j] = __ldg(Wt + (size_t)j * VP + v1); }
    float b0[3], b1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { b0[c] = __ldg(base + c * VP + v0); b1[c] = __ldg(base + c * VP + v1); }
    float jx0[NXR > 0 ? NXR : 1], jx1[NXR > 0 ? NXR : 1];
#pragma unroll
    for (int r = 0; r < NXR; ++r) {
      jx0[r] = r < nx ? __ldg(Jx + (size_t)r * VP + v0) : 0.f;
      jx1[r] = r < nx ? __ldg(Jx + (size_t)r * VP + v1) : 0.f;
    }
    const int s0 = slot ? slot[v0] : -1, s1 = slot ? slot[v1] : -1;
#pragma unroll
    for (int f = 0; f < FB; ++f) {
      const int n = n0 + f;
      float* g = sG + (f & 1) * SK_TILE * SG_STRIDE;
      // ---- part 1
      float dv0[3] = {0.f, 0.f, 0.f}, dv1[3] = {0.f, 0.f, 0.f};
      float p0[3] = {b0[0], b0[1], b0[2]}, p1[3] = {b1[0], b1[1], b1[2]};
      if (n < N) {
        if (PO) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            p0[c] += __ldcs(PO + ((size_t)n * 3 + c) * VP + v0);
            p1[c] += __ldcs(PO + ((size_t)n * 3 + c) * VP + v1);
          }
        }
        if (dV) {
          if (v0 < NV) {
            const float* s = dV + ((size_t)n * NV + v0) * 3;
            dv0[0] = __ldcs(s); dv0[1] = __ldcs(s + 1); dv0[2] = __ldcs(s + 2);
          }
          if (v1 < NV) {
            const float* s = dV + ((size_t)n * NV + v1) * 3;
            dv1[0] = __ldcs(s); dv1[1] = __ldcs(s + 1); dv1[2] = __ldcs(s + 2);
          }
        }
        if (signs) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const uint32_t* s = signs + ((size_t)n * 3 + c) * 2 * (VP / 32);
            int wd0 = t * 8 + warp, wd1 = wd0 + 4;
            uint32_t pa = __ldg(s + wd0), na = __ldg(s + (VP / 32) + wd0);
            uint32_t pb = __ldg(s + wd1), nb_ = __ldg(s + (VP / 32) + wd1);
            dv0[c] += sign_scale * ((float)((pa >> lane) & 1u) - (float)((na >> lane) & 1u));
            dv1[c] += sign_scale * ((float)((pb >> lane) & 1u) - (float)((nb_ >> lane) & 1u));
          }
        }
        const float* dj = sdJ + f * 96;
        if (s0 >= 0) { dv0[0] += dj[s0 * 3]; dv0[1] += dj[s0 * 3 + 1]; dv0[2] += dj[s0 * 3 + 2]; }
        if (s1 >= 0) { dv1[0] += dj[s1 * 3]; dv1[1] += dj[s1 * 3 + 1]; dv1[2] += dj[s1 * 3 + 2]; }
#pragma unroll
        for (int r = 0; r < NXR; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float d = dj[(np + r) * 3 + c];
            dv0[c] += jx0[r] * d;
            dv1[c] += jx1[r] * d;
          }
        if (dPO) {
          float R0[9], R1[9];
          accum_R2(sA + f * 288, w0, w1, R0, R1);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            __stcs(dPO + ((size_t)n * 3 + c) * VP + v0, R0[c] * dv0[0] + R0[3 + c] * dv0[1] + R0[6 + c] * dv0[2]);
            __stcs(dPO + ((size_t)n * 3 + c) * VP + v1, R1[c] * dv1[0] + R1[3 + c] * dv1[1] + R1[6 + c] * dv1[2]);
          }
        }
      }
      {
        // G[r*4 + c4] = dv[r] * [p;1][c4], stored as 4 groups of 3 (+ pad): comps {0,1,2},{3,4,5},{6,7,8},{9,10,11}
        float G0[12], G1[12];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          G0[4 * r] = dv0[r] * p0[0]; G0[4 * r + 1] = dv0[r] * p0[1]; G0[4 * r + 2] = dv0[r] * p0[2]; G0[4 * r + 3] = dv0[r];
          G1[4 * r] = dv1[r] * p1[0]; G1[4 * r + 1] = dv1[r] * p1[1]; G1[4 * r + 2] = dv1[r] * p1[2]; G1[4 * r + 3] = dv1[r];
        }
        float4* g0 = reinterpret_cast<float4*>(g + tid * SG_STRIDE);
        float4* g1 = reinterpret_cast<float4*>(g + (tid + SK_THREADS) * SG_STRIDE);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          g0[q] = make_float4(G0[3 * q], G0[3 * q + 1], G0[3 * q + 2], 0.f);
          g1[q] = make_float4(G1[3 * q], G1[3 * q + 1], G1[3 * q + 2], 0.f);
        }
      }
      __syncthreads();
      // ---- part 2: this warp's 64 vertices of the tile
      {
        const float* wbase = sW2 + (warp * 64) * 32 + jt * 4;
        const float* gbase = g + (warp * 64) * SG_STRIDE + cq * 4;
#pragma unroll 8
        for (int vv = 0; vv < 64; ++vv) {
          float4 wq = *reinterpret_cast<const float4*>(wbase + vv * 32);
          float4 gq = *reinterpret_cast<const float4*>(gbase + vv * SG_STRIDE);
          acc[f][0] += wq.x * gq.x; acc[f][1] += wq.x * gq.y; acc[f][2] += wq.x * gq.z;
          acc[f][3] += wq.y * gq.x; acc[f][4] += wq.y * gq.y; acc[f][5] += wq.y * gq.z;
          acc[f][6] += wq.z * gq.x; acc[f][7] += wq.z * gq.y; acc[f][8] += wq.z * gq.z;
        }
      }
    }
  }
  // ---- cross-warp reduction of dA and write-out
  __syncthreads();
  float* red = sW2;  // 4 * FB * 288 floats <= 256*32 for FB <= 7; for FB == 8 spills into sG (contiguous)
#pragma unroll
  for (int f = 0; f < FB; ++f)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) red[(warp * FB + f) * 288 + (3 * jt + a) * 12 + 3 * cq + b] = acc[f][3 * a + b];
  __syncthreads();
  for (int i = tid; i < FB * 288; i += SK_THREADS) {
    int f = i / 288;
    if (n0 + f < N)
      dA[(size_t)(n0 + f) * 288 + (i % 288)] = red[i] + red[FB * 288 + i] + red[2 * FB * 288 + i] + red[3 * FB * 288 + i];
  }
}

template <int FB>
constexpr size_t bwd_smem_bytes() {
  return sizeof(float) * (FB * 288 + FB * 96 + SK_TILE * 32 + 2 * SK_TILE * SG_STRIDE);
}

}  // namespace

constexpr int FWD_FB = 8;
constexpr int BWD_FB = 4;
constexpr int L1_FB = 4;

extern "C" int nemo_b200_skin_fwd(const float* A, const float* PO, const float* base, const float* Wt, const float* Jx,
                                  int nx, const signed char* slot, int np, float* V, float* JX, int N, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(A && base && Wt && N >= 0 && nx >= 0 && nx <= 10 && np >= 0 && np <= 22, "skin_fwd: bad args");
  NB_REQUIRE((nx == 0 || Jx) && (np == 0 || slot) && (nx + np == 0 || JX), "skin_fwd: joint read-out pointers missing");
  dim3 grid(cdiv(N, FWD_FB));
  cudaStream_t s = S(stream);
#define NB_SKF(NXP, WV) k_skin_fwd<FWD_FB, NXP, WV><<<grid, SK_THREADS, 0, s>>>(A, PO, base, Wt, Jx, nx, np ? slot : nullptr, np, V, JX, N)
  if (V) {
    if (nx == 0) NB_SKF(0, true); else if (nx == 1) NB_SKF(4, true); else NB_SKF(32, true);
  } else {
    if (nx == 0) NB_SKF(0, false); else if (nx == 1) NB_SKF(4, false); else NB_SKF(32, false);
  }
#undef NB_SKF
  NB_LAUNCH_CHECK("skin_fwd");
}

extern "C" int nemo_b200_skin_l1_fwd(const float* A_a, const float* PO_a, const float* A_b, const float* PO_b,
                                     const float* base, const float* Wt, uint32_t* signs, float* l1_sum, int N,
                                     void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(A_a && PO_a && A_b && PO_b && base && Wt && l1_sum && N >= 0, "skin_l1_fwd: bad args");
  k_skin_l1_fwd<L1_FB><<<cdiv(N, L1_FB), SK_THREADS, 0, S(stream)>>>(A_a, PO_a, A_b, PO_b, base, Wt, signs, l1_sum, N);
  NB_LAUNCH_CHECK("skin_l1_fwd");
}

extern "C" int nemo_b200_skin_bwd(const float* A, const float* PO, const float* base, const float* Wt, const float* W2,
                                  const float* Jx, int nx, const signed char* slot, int np, const float* dV,
                                  const float* dJX, const uint32_t* signs, float sign_scale, float* dPO, float* dA,
                                  int N, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(A && base && Wt && W2 && dA && N >= 0 && nx >= 0 && nx <= 10 && np >= 0 && np <= 22,
             "skin_bwd: bad args");
  NB_REQUIRE(!dJX || ((nx == 0 || Jx) && (np == 0 || slot)), "skin_bwd: joint read-out pointers missing");
  constexpr size_t smem = bwd_smem_bytes<BWD_FB>();
  static_assert(4 * BWD_FB * 288 <= SK_TILE * 32 + 2 * SK_TILE * SG_STRIDE, "reduction scratch too small");
  const int nxe = dJX ? nx : 0;
  const signed char* sl = (dJX && np) ? slot : nullptr;
  cudaStream_t s = S(stream);
  dim3 grid(cdiv(N, BWD_FB));
#define NB_SKB(NXR)                                                                                              \
  do {                                                                                                           \
    NB_CUDA(cudaFuncSetAttribute(k_skin_bwd<BWD_FB, NXR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_skin_bwd<BWD_FB, NXR><<<grid, SK_THREADS, smem, s>>>(A, PO, base, Wt, W2, Jx, nxe, sl, np, dV, dJX, signs,  \
                                                           sign_scale, dPO, dA, N);                              \
  } while (0)
  if (nxe == 0) NB_SKB(0); else if (nxe == 1) NB_SKB(1); else NB_SKB(10);
#undef NB_SKB
  NB_LAUNCH_CHECK("skin_bwd");
}
