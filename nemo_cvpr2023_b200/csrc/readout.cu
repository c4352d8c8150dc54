// Joint read-outs of the skinned mesh without touching the mesh.
// Reference arithmetic: smplx VertexJointSelector (vertex picks), hmr/smpl.py:33-36
// (joints_extra = J_regressor_extra . vertices), on top of lbs.py:236-252.
//
// A read-out q is a fixed linear functional of the vertices, joint_q = sum_v w_q[v] * v[n,v] (w_q one-hot for a
// picked vertex, a J_regressor_extra row for a regressed joint).  Skinning is linear in [v_posed;1] per joint,
//   v[n,v] = sum_j W[v,j] (A[n,j]_3x3 . p[n,v] + A[n,j]_t),     p[n,v] = base[v] + F[n,:] . posedirs[:, v]
// so the sum over vertices can be taken BEFORE the per-frame transforms are applied:
//   joint_q[n] = sum_j A[n,j]_3x3 . z[n,q,j] + A[n,j]_t * s[q,j]
//   z[n,q,j]   = sum_v w_q[v] W[v,j] p[n,v] = Z0[q,j] + F[n,:] . MZ[:, (q,j)]     s[q,j] = sum_v w_q[v] W[v,j]
// with MZ = posedirs contracted with w_q (x) W once per asset set (host, float64).  Z (N, ldz) is produced by the
// tensor-core GEMM; the kernels here are the tiny per-frame parts.  For a picked vertex z collapses to its own
// v_posed (3 values) and the 24 weights of that vertex.  Columns of Z: [np x 3 | nx x 24 x 3].
#include "common.cuh"

using namespace nb;

namespace {

__global__ void __launch_bounds__(128) k_readout_fwd(const float* __restrict__ A, const float* __restrict__ Z, int ldz,
                                                     const float* __restrict__ Wp, int np, const float* __restrict__ Sj,
                                                     int nx, float* __restrict__ JX, int N) {
  const int nq = np + nx;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * nq) return;
  const int q = (int)(i % nq), n = (int)(i / nq);
  const float* An = A + (size_t)n * 288;
  const float* z = Z + (size_t)n * ldz;
  float o0 = 0.f, o1 = 0.f, o2 = 0.f;
  if (q < np) {
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = 0.f;
    const float* w = Wp + q * NJ;
    for (int j = 0; j < NJ; ++j) {
      const float wj = __ldg(w + j);
#pragma unroll
      for (int k = 0; k < 12; ++k) T[k] = fmaf(wj, __ldg(An + j * 12 + k), T[k]);
    }
    const float p0 = z[3 * q], p1 = z[3 * q + 1], p2 = z[3 * q + 2];
    o0 = T[0] * p0 + T[1] * p1 + T[2] * p2 + T[3];
    o1 = T[4] * p0 + T[5] * p1 + T[6] * p2 + T[7];
    o2 = T[8] * p0 + T[9] * p1 + T[10] * p2 + T[11];
  } else {
    const int r = q - np;
    const float* zr = z + 3 * np + r * 72;
    const float* s = Sj + r * NJ;
    for (int j = 0; j < NJ; ++j) {
      const float* a = An + j * 12;
      const float z0 = zr[3 * j], z1 = zr[3 * j + 1], z2 = zr[3 * j + 2], sj = __ldg(s + j);
      o0 += __ldg(a + 0) * z0 + __ldg(a + 1) * z1 + __ldg(a + 2) * z2 + __ldg(a + 3) * sj;
      o1 += __ldg(a + 4) * z0 + __ldg(a + 5) * z1 + __ldg(a + 6) * z2 + __ldg(a + 7) * sj;
      o2 += __ldg(a + 8) * z0 + __ldg(a + 9) * z1 + __ldg(a + 10) * z2 + __ldg(a + 11) * sj;
    }
  }
  float* d = JX + (size_t)i * 3;
  d[0] = o0; d[1] = o1; d[2] = o2;
}

// dA[n,j] (3x4) = sum_q  (pick)      Wp[q,j] * dJX[n,q] (x) [z_q; 1]
//                      + (regressed) dJX[n,r] (x) [z[r,j]; s[r,j]]          one thread per (n, j)
__global__ void __launch_bounds__(96) k_readout_bwd_dA(const float* __restrict__ Z, int ldz, const float* __restrict__ Wp,
                                                       int np, const float* __restrict__ Sj, int nx,
                                                       const float* __restrict__ dJX, float* __restrict__ dA,
                                                       int accumulate, int N) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * NJ) return;
  const int j = (int)(i % NJ), n = (int)(i / NJ);
  const int nq = np + nx;
  const float* z = Z + (size_t)n * ldz;
  const float* g = dJX + (size_t)n * nq * 3;
  float d[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) d[k] = 0.f;
  for (int q = 0; q < np; ++q) {
    const float w = __ldg(Wp + q * NJ + j);
    const float ph[4] = {z[3 * q], z[3 * q + 1], z[3 * q + 2], 1.f};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float gw = w * g[3 * q + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) d[4 * r + c] = fmaf(gw, ph[c], d[4 * r + c]);
    }
  }
  for (int r2 = 0; r2 < nx; ++r2) {
    const float* zr = z + 3 * np + r2 * 72 + 3 * j;
    const float ph[4] = {zr[0], zr[1], zr[2], __ldg(Sj + r2 * NJ + j)};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float gg = g[3 * (np + r2) + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) d[4 * r + c] = fmaf(gg, ph[c], d[4 * r + c]);
    }
  }
  float* o = dA + (size_t)i * 12;
#pragma unroll
  for (int k = 0; k < 12; ++k) o[k] = accumulate ? o[k] + d[k] : d[k];
}

// dZ: pick q -> (sum_j Wp[q,j] A[n,j]_3x3)^T dJX[n,q];  regressed (r,j) -> A[n,j]_3x3^T dJX[n,r].
// one thread per (n, column triple); pad columns of the row are zeroed by the first thread of the frame
__global__ void __launch_bounds__(128) k_readout_bwd_dZ(const float* __restrict__ A, const float* __restrict__ Wp, int np,
                                                        int nx, const float* __restrict__ dJX, float* __restrict__ dZ,
                                                        int ldz, int N) {
  const int nt = np + nx * NJ;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * nt) return;
  const int t = (int)(i % nt), n = (int)(i / nt);
  const int nq = np + nx;
  const float* An = A + (size_t)n * 288;
  const float* g = dJX + (size_t)n * nq * 3;
  float* o = dZ + (size_t)n * ldz;
  if (t == 0)
    for (int k = 3 * nt; k < ldz; ++k) o[k] = 0.f;
  float r0 = 0.f, r1 = 0.f, r2 = 0.f;
  if (t < np) {
    const float g0 = g[3 * t], g1 = g[3 * t + 1], g2 = g[3 * t + 2];
    for (int j = 0; j < NJ; ++j) {
      const float w = __ldg(Wp + t * NJ + j);
      const float* a = An + j * 12;
      r0 = fmaf(w, __ldg(a + 0) * g0 + __ldg(a + 4) * g1 + __ldg(a + 8) * g2, r0);
      r1 = fmaf(w, __ldg(a + 1) * g0 + __ldg(a + 5) * g1 + __ldg(a + 9) * g2, r1);
      r2 = fmaf(w, __ldg(a + 2) * g0 + __ldg(a + 6) * g1 + __ldg(a + 10) * g2, r2);
    }
  } else {
    const int rr = (t - np) / NJ, j = (t - np) % NJ;
    const float g0 = g[3 * (np + rr)], g1 = g[3 * (np + rr) + 1], g2 = g[3 * (np + rr) + 2];
    const float* a = An + j * 12;
    r0 = __ldg(a + 0) * g0 + __ldg(a + 4) * g1 + __ldg(a + 8) * g2;
    r1 = __ldg(a + 1) * g0 + __ldg(a + 5) * g1 + __ldg(a + 9) * g2;
    r2 = __ldg(a + 2) * g0 + __ldg(a + 6) * g1 + __ldg(a + 10) * g2;
  }
  o[3 * t] = r0; o[3 * t + 1] = r1; o[3 * t + 2] = r2;
}

}  // namespace

extern "C" int nemo_b200_readout_fwd(const float* A, const float* Z, int ldz, const float* Wp, int np, const float* Sj,
                                     int nx, float* JX, int N, void* stream) {
  if (N == 0 || np + nx == 0) return 0;
  NB_REQUIRE(A && Z && JX && N >= 0 && np >= 0 && nx >= 0 && (np == 0 || Wp) && (nx == 0 || Sj) &&
                 ldz >= 3 * np + 72 * nx, "readout_fwd: bad args");
  const long long total = (long long)N * (np + nx);
  k_readout_fwd<<<cdiv(total, 128), 128, 0, S(stream)>>>(A, Z, ldz, Wp, np, Sj, nx, JX, N);
  NB_LAUNCH_CHECK("readout_fwd");
}

extern "C" int nemo_b200_readout_bwd(const float* A, const float* Z, int ldz, const float* Wp, int np, const float* Sj,
                                     int nx, const float* dJX, float* dA, int accumulate, float* dZ, int N,
                                     void* stream) {
  if (N == 0 || np + nx == 0) return 0;
  NB_REQUIRE(A && Z && dJX && dA && dZ && N >= 0 && np >= 0 && nx >= 0 && (np == 0 || Wp) && (nx == 0 || Sj) &&
                 ldz >= 3 * np + 72 * nx, "readout_bwd: bad args");
  k_readout_bwd_dA<<<cdiv((long long)N * NJ, 96), 96, 0, S(stream)>>>(Z, ldz, Wp, np, Sj, nx, dJX, dA, accumulate, N);
  NB_LAUNCH_CHECK_CONT("readout_bwd_dA");
  const long long total = (long long)N * (np + nx * NJ);
  k_readout_bwd_dZ<<<cdiv(total, 128), 128, 0, S(stream)>>>(A, Wp, np, nx, dJX, dZ, ldz, N);
  NB_LAUNCH_CHECK("readout_bwd_dZ");
}
