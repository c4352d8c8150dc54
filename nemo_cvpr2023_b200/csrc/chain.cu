// K6: the 24-joint kinematic chain, forward and backward, one WARP per frame (lane = joint).
// Reference: batch_rigid_transform + transform_mat, human_body_prior/body_model/lbs.py:337-404; the pose feature
// (R[1:] - I) of lbs.py:221-228 and the quaternion / skew Rodrigues of the two extra vposer_loss meshes
// (hmr/geometry.py:9-45, lbs.py:303-334) are fused into the forward so a mesh costs one small launch, not four.
//
// Lane j keeps the 3x4 world transform G_j of joint j in registers.  The tree is walked by DEPTH (SMPL: 9 levels, not
// 24 steps): at level d every joint of that depth reads its parent's transform from a per-warp shared-memory table
// (three 16-byte loads) and composes G_j = G_p . [R_j | J_j - J_p].  The backward recomputes that sweep, then walks the
// levels upwards: a joint whose gradient is complete forms dR_j = G_p^T dG_j, dt_j, and a 15-float message for its
// parent (dG_p += dG_j . [R_j | t_j]^T, dJ_p -= dt_j) that the parent lane collects over its children mask.  Any tree
// given by `parents` works (parent < 0 = root).  4096 frames = 4096 warps instead of the 512 of the four-lanes-per-
// frame version with its 24 dependent local-memory steps.
#include "common.cuh"
#include "rotmath.cuh"

using namespace nb;

namespace {

constexpr int CW = 8;  // warps per CTA = frames in flight per CTA
constexpr unsigned FULL = 0xffffffffu;

struct Tree {
  int parent, depth, maxdepth;
  unsigned children;
};

__device__ __forceinline__ Tree tree_of(const int* sp, int lane) {
  Tree t;
  t.parent = -1;
  t.depth = 0;
  if (lane < NJ) {
    t.parent = lane == 0 ? -1 : sp[lane];
    int q = lane, d = 0;
    while (q > 0 && sp[q] >= 0 && d < NJ) {
      q = sp[q];
      ++d;
    }
    t.depth = d;
  }
  t.children = 0u;
#pragma unroll 1
  for (int j = 0; j < NJ; ++j) {
    const unsigned m = __ballot_sync(FULL, lane < NJ && t.parent == j);
    if (lane == j) t.children = m;
  }
  int md = t.depth;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) md = max(md, __shfl_xor_sync(FULL, md, o));
  t.maxdepth = md;
  return t;
}

__device__ __forceinline__ void lds12(const float* p, float (&g)[12]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4),
               c = *reinterpret_cast<const float4*>(p + 8);
  g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
  g[8] = c.x; g[9] = c.y; g[10] = c.z; g[11] = c.w;
}
__device__ __forceinline__ void sts12(float* p, const float (&g)[12]) {
  *reinterpret_cast<float4*>(p) = make_float4(g[0], g[1], g[2], g[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(g[4], g[5], g[6], g[7]);
  *reinterpret_cast<float4*>(p + 8) = make_float4(g[8], g[9], g[10], g[11]);
}

// world transforms of one frame: G (own, row-major 3x4: G[4r + c], c = 3 the translation) and Gp (the parent's).
// sG: this warp's [NJ][12] table.  R: own rotation (row-major), J / Jp: own and parent's rest joint.
__device__ __forceinline__ void sweep_down(const Tree& t, int lane, float* sG, const float (&R)[9], const float (&J)[3],
                                           const float (&tr)[3], float (&G)[12], float (&Gp)[12]) {
  if (lane < NJ && t.parent < 0) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      G[4 * r] = R[3 * r]; G[4 * r + 1] = R[3 * r + 1]; G[4 * r + 2] = R[3 * r + 2]; G[4 * r + 3] = J[r];
    }
    sts12(sG + 12 * lane, G);
  }
#pragma unroll 1
  for (int d = 1; d <= t.maxdepth; ++d) {
    __syncwarp();
    if (lane < NJ && t.depth == d && t.parent >= 0) {
      lds12(sG + 12 * t.parent, Gp);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float g0 = Gp[4 * r], g1 = Gp[4 * r + 1], g2 = Gp[4 * r + 2];
        G[4 * r] = g0 * R[0] + g1 * R[3] + g2 * R[6];
        G[4 * r + 1] = g0 * R[1] + g1 * R[4] + g2 * R[7];
        G[4 * r + 2] = g0 * R[2] + g1 * R[5] + g2 * R[8];
        G[4 * r + 3] = g0 * tr[0] + g1 * tr[1] + g2 * tr[2] + Gp[4 * r + 3];
      }
      sts12(sG + 12 * lane, G);
    }
  }
}

// this frame's rotations into the warp's shared-memory slot sR[216] and the lane's registers
__device__ __forceinline__ void load_rotations(const float* __restrict__ R, const float* __restrict__ aa, int lda, int kind,
                                               const float* __restrict__ R0, size_t n, int lane, float* sR, float (&Rj)[9]) {
  if (aa == nullptr) {
    for (int k = lane; k < 216; k += 32) sR[k] = __ldg(R + n * 216 + k);
    __syncwarp();
    if (lane < NJ) {
#pragma unroll
      for (int e = 0; e < 9; ++e) Rj[e] = sR[9 * lane + e];
    }
  } else {
    if (lane == 0) {
#pragma unroll
      for (int e = 0; e < 9; ++e) Rj[e] = __ldg(R0 + n * 216 + e);
    } else if (lane < NJ) {
      const float* a = aa + n * lda + 3 * (lane - 1);
      const float th[3] = {__ldg(a), __ldg(a + 1), __ldg(a + 2)};
      if (kind == 0) rodrigues_quat_fwd(th, Rj);
      else rodrigues_skew_fwd(th, Rj);
    }
    if (lane < NJ) {
#pragma unroll
      for (int e = 0; e < 9; ++e) sR[9 * lane + e] = Rj[e];
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(CW * 32) k_chain_fwd(const float* __restrict__ R, const float* __restrict__ aa, int lda,
                                                       int kind, const float* __restrict__ R0, const float* __restrict__ J,
                                                       int js, const int* __restrict__ parents, float* __restrict__ Rout,
                                                       float* __restrict__ F, int ldf, float* __restrict__ A,
                                                       float* __restrict__ JP, int N) {
  __shared__ int sp[NJ];
  __shared__ __align__(16) float sG[CW][NJ * 12];
  __shared__ float sRm[CW][216];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < NJ) sp[threadIdx.x] = parents[threadIdx.x];
  __syncthreads();
  const Tree t = tree_of(sp, lane);
  float* sR = sRm[warp];
  for (size_t n = (size_t)blockIdx.x * CW + warp; n < (size_t)N; n += (size_t)gridDim.x * CW) {
    float Rj[9], Jj[3] = {0.f, 0.f, 0.f}, tr[3] = {0.f, 0.f, 0.f}, G[12], Gp[12];
    load_rotations(R, aa, lda, kind, R0, n, lane, sR, Rj);
    if (aa != nullptr && Rout != nullptr)
      for (int k = lane; k < 216; k += 32) Rout[n * 216 + k] = sR[k];
    if (F != nullptr)
      for (int k = lane; k < 207; k += 32) {
        const int e = k % 9;
        F[n * ldf + k] = sR[9 + k] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
      }
    if (lane < NJ) {
      const float* Jn = J + n * js;
#pragma unroll
      for (int c = 0; c < 3; ++c) Jj[c] = __ldg(Jn + 3 * lane + c);
      if (t.parent >= 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) tr[c] = Jj[c] - __ldg(Jn + 3 * t.parent + c);
      }
    }
    sweep_down(t, lane, sG[warp], Rj, Jj, tr, G, Gp);
    if (lane < NJ) {
      float* An = A + n * 288 + 12 * lane;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float a3 = G[4 * r + 3] - (G[4 * r] * Jj[0] + G[4 * r + 1] * Jj[1] + G[4 * r + 2] * Jj[2]);
        *reinterpret_cast<float4*>(An + 4 * r) = make_float4(G[4 * r], G[4 * r + 1], G[4 * r + 2], a3);
        JP[n * 72 + 3 * lane + r] = G[4 * r + 3];
      }
    }
    __syncwarp();  // the tables are rewritten by the next frame
  }
}

__global__ void __launch_bounds__(CW * 32) k_chain_bwd(const float* __restrict__ R, const float* __restrict__ J, int js,
                                                       const int* __restrict__ parents, const float* __restrict__ dA,
                                                       const float* __restrict__ dAt, const float* __restrict__ dJP,
                                                       const float* __restrict__ dF, int ldf,
                                                       float* __restrict__ dR, float* __restrict__ dJ, int N) {
  __shared__ int sp[NJ];
  __shared__ __align__(16) float sG[CW][NJ * 12];
  __shared__ __align__(16) float sM[CW][NJ * 16];
  __shared__ float sRm[CW][216];
  __shared__ float sdj[CW][NJ * 3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < NJ) sp[threadIdx.x] = parents[threadIdx.x];
  __syncthreads();
  const Tree t = tree_of(sp, lane);
  float* sR = sRm[warp];
  float* sMsg = sM[warp];
  float djacc[3] = {0.f, 0.f, 0.f};
  for (size_t n = (size_t)blockIdx.x * CW + warp; n < (size_t)N; n += (size_t)gridDim.x * CW) {
    float Rj[9], Jj[3] = {0.f, 0.f, 0.f}, tr[3] = {0.f, 0.f, 0.f}, G[12], Gp[12];
    load_rotations(R, nullptr, 0, 0, nullptr, n, lane, sR, Rj);
    float dG[12], dJj[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 12; ++k) dG[k] = 0.f;
    if (lane < NJ) {
      const float* Jn = J + n * js;
#pragma unroll
      for (int c = 0; c < 3; ++c) Jj[c] = __ldg(Jn + 3 * lane + c);
      if (t.parent >= 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) tr[c] = Jj[c] - __ldg(Jn + 3 * t.parent + c);
      }
    }
    sweep_down(t, lane, sG[warp], Rj, Jj, tr, G, Gp);
    if (lane < NJ) {
      // upstream: A_j = [G.R | G.t - G.R J_j], JP_j = G.t
      float a[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) a[k] = 0.f;
      if (dA) {
        const float4* p = reinterpret_cast<const float4*>(dA + n * 288 + 12 * lane);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float4 v = __ldg(p + r);
          a[4 * r] = v.x; a[4 * r + 1] = v.y; a[4 * r + 2] = v.z; a[4 * r + 3] = v.w;
        }
      }
      if (dAt) {
#pragma unroll
        for (int k = 0; k < 12; ++k) a[k] += __ldg(dAt + n * 288 + k * NJ + lane);
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float a3 = a[4 * r + 3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          dG[4 * r + c] = a[4 * r + c] - a3 * Jj[c];
          dJj[c] -= G[4 * r + c] * a3;
        }
        dG[4 * r + 3] = a3 + (dJP ? __ldg(dJP + n * 72 + 3 * lane + r) : 0.f);
      }
    }
    __syncwarp();  // everybody is done with sR (rotations): it now collects dR
    // ---- up the tree
#pragma unroll 1
    for (int d = t.maxdepth; d >= 1; --d) {
      if (lane < NJ && t.depth == d && t.parent >= 0) {
        float m[16];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float d0 = dG[4 * r], d1 = dG[4 * r + 1], d2 = dG[4 * r + 2], d3 = dG[4 * r + 3];
          m[4 * r] = d0 * Rj[0] + d1 * Rj[1] + d2 * Rj[2] + d3 * tr[0];
          m[4 * r + 1] = d0 * Rj[3] + d1 * Rj[4] + d2 * Rj[5] + d3 * tr[1];
          m[4 * r + 2] = d0 * Rj[6] + d1 * Rj[7] + d2 * Rj[8] + d3 * tr[2];
          m[4 * r + 3] = d3;
        }
        // dR_j[a][b] = sum_r Gp[r][a] dG[r][b], dt[a] = sum_r Gp[r][a] dG[r][3]
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float g0 = Gp[a], g1 = Gp[4 + a], g2 = Gp[8 + a];
#pragma unroll
          for (int b = 0; b < 3; ++b) sR[9 * lane + 3 * a + b] = g0 * dG[b] + g1 * dG[4 + b] + g2 * dG[8 + b];
          const float dt = g0 * dG[3] + g1 * dG[7] + g2 * dG[11];
          m[12 + a] = dt;
          dJj[a] += dt;
        }
        m[15] = 0.f;
        float4* q = reinterpret_cast<float4*>(sMsg + 16 * lane);
        q[0] = make_float4(m[0], m[1], m[2], m[3]);
        q[1] = make_float4(m[4], m[5], m[6], m[7]);
        q[2] = make_float4(m[8], m[9], m[10], m[11]);
        q[3] = make_float4(m[12], m[13], m[14], m[15]);
      }
      __syncwarp();
      if (lane < NJ && t.depth == d - 1) {
        unsigned ch = t.children;
        while (ch) {
          const int i = __ffs(ch) - 1;
          ch &= ch - 1;
          const float4* q = reinterpret_cast<const float4*>(sMsg + 16 * i);
          const float4 m0 = q[0], m1 = q[1], m2 = q[2], m3 = q[3];
          dG[0] += m0.x; dG[1] += m0.y; dG[2] += m0.z; dG[3] += m0.w;
          dG[4] += m1.x; dG[5] += m1.y; dG[6] += m1.z; dG[7] += m1.w;
          dG[8] += m2.x; dG[9] += m2.y; dG[10] += m2.z; dG[11] += m2.w;
          dJj[0] -= m3.x; dJj[1] -= m3.y; dJj[2] -= m3.z;
        }
      }
      __syncwarp();
    }
    if (lane < NJ && t.parent < 0) {  // root(s): G = [R | J]
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        sR[9 * lane + 3 * r] = dG[4 * r]; sR[9 * lane + 3 * r + 1] = dG[4 * r + 1]; sR[9 * lane + 3 * r + 2] = dG[4 * r + 2];
        dJj[r] += dG[4 * r + 3];
      }
    }
    __syncwarp();
    // + the gradient of the pose feature (R[1:] - I), which reaches the rotations of joints 1..23 directly
    for (int k = lane; k < 216; k += 32) dR[n * 216 + k] = sR[k] + ((dF != nullptr && k >= 9) ? __ldg(dF + n * ldf + (k - 9)) : 0.f);
    if (dJ != nullptr && lane < NJ) {
      if (js != 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) dJ[n * 72 + 3 * lane + c] = dJj[c];
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) djacc[c] += dJj[c];
      }
    }
    __syncwarp();
  }
  if (dJ != nullptr && js == 0) {  // shared rest joints: sum over the frames of the CTA, one atomic per value
    if (lane < NJ) {
#pragma unroll
      for (int c = 0; c < 3; ++c) sdj[warp][3 * lane + c] = djacc[c];
    }
    __syncthreads();
    if (threadIdx.x < NJ * 3) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < CW; ++w) v += sdj[w][threadIdx.x];
      atomicAdd(dJ + threadIdx.x, v);
    }
  }
}

int chain_grid(int N) { return min(cdiv(N, CW), 4 * num_sms()); }

}  // namespace

extern "C" int nemo_b200_pose_chain_fwd(const float* R, const float* aa, int lda, int kind, const float* R0, const float* J,
                                        int j_stride, const int* parents, float* R_out, float* F, int ldf, float* A,
                                        float* JP, int N, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(J && parents && A && JP && N >= 0 && (j_stride == 0 || j_stride == 72), "pose_chain_fwd: bad args");
  NB_REQUIRE((R != nullptr) != (aa != nullptr), "pose_chain_fwd: exactly one of R and aa");
  NB_REQUIRE(aa == nullptr || (R0 && lda >= 69 && (kind == 0 || kind == 1)), "pose_chain_fwd: aa needs R0, lda >= 69, kind 0 / 1");
  NB_REQUIRE(F == nullptr || ldf >= 207, "pose_chain_fwd: ldf >= 207");
  k_chain_fwd<<<chain_grid(N), CW * 32, 0, S(stream)>>>(R, aa, lda, kind, R0, J, j_stride, parents, R_out, F, ldf, A, JP, N);
  NB_LAUNCH_CHECK("pose_chain_fwd");
}

extern "C" int nemo_b200_chain_fwd(const float* R, const float* J, int j_stride, const int* parents, float* A, float* JP,
                                   int N, void* stream) {
  NB_REQUIRE(R, "chain_fwd: bad args");
  return nemo_b200_pose_chain_fwd(R, nullptr, 0, 0, nullptr, J, j_stride, parents, nullptr, nullptr, 0, A, JP, N, stream);
}

extern "C" int nemo_b200_chain_bwd2(const float* R, const float* J, int j_stride, const int* parents, const float* dA,
                                    const float* dAt, const float* dJP, const float* dF, int ldf, float* dR, float* dJ,
                                    int N, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(R && J && parents && dR && N >= 0 && (j_stride == 0 || j_stride == 72), "chain_bwd: bad args");
  NB_REQUIRE(dF == nullptr || ldf >= 207, "chain_bwd: ldf >= 207");
  // shared joints accumulate with one atomic per value and CTA: fewer, longer CTAs
  const int grid = (dJ && j_stride == 0) ? min(cdiv(N, CW), num_sms()) : chain_grid(N);
  k_chain_bwd<<<grid, CW * 32, 0, S(stream)>>>(R, J, j_stride, parents, dA, dAt, dJP, dF, ldf, dR, dJ, N);
  NB_LAUNCH_CHECK("chain_bwd");
}

extern "C" int nemo_b200_chain_bwd(const float* R, const float* J, int j_stride, const int* parents, const float* dA,
                                   const float* dAt, const float* dJP, float* dR, float* dJ, int N, void* stream) {
  return nemo_b200_chain_bwd2(R, J, j_stride, parents, dA, dAt, dJP, nullptr, 0, dR, dJ, N, stream);
}
