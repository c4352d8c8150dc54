// K1 (motion-field input), K3 (rotations), K6 (kinematic chain), K7b (projection + keypoint loss),
// K8 (GMM prior), VPoser KL, fused Adam: the latency-bound, per-frame part of the fit path.
#include <cstdlib>
#include "pathmath.cuh"
#include "rotmath.cuh"

using namespace nb;

// =============================================================================== rotations
// Strided addressing shared by all rotation kernels: rotation i lives in row n = i / per_row at slot
// j = i % per_row; a tensor with row stride ld and w floats per rotation is read at base + n*ld + w*j.
// (contiguous (B,w): per_row = 1, ld = w.)  This lets the kernels work in place on slices such as
// rot6d = Y[:, :144] of a (N,147) MLP output or poses = aa[:, 3:72].
#define NB_AT(base, ld, w) ((base) + (size_t)(i / per_row) * (ld) + (w) * (i % per_row))

__global__ void k_rot6d_fwd(const float* __restrict__ x6, int ldx, float* __restrict__ R, int ldr, int per_row, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float x[6], r[9];
  const float* xi = NB_AT(x6, ldx, 6);
#pragma unroll
  for (int k = 0; k < 6; ++k) x[k] = xi[k];
  rot6d_fwd(x, r);
  float* ro = NB_AT(R, ldr, 9);
#pragma unroll
  for (int k = 0; k < 9; ++k) ro[k] = r[k];
}
__global__ void k_rot6d_bwd(const float* __restrict__ x6, int ldx, const float* __restrict__ dR, int ldr,
                            float* __restrict__ dx6, int ldd, int per_row, int acc, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float x[6], g[9], dx[6];
  const float* xi = NB_AT(x6, ldx, 6);
  const float* gi = NB_AT(dR, ldr, 9);
#pragma unroll
  for (int k = 0; k < 6; ++k) x[k] = xi[k];
#pragma unroll
  for (int k = 0; k < 9; ++k) g[k] = gi[k];
  rot6d_bwd(x, g, dx);
  float* o = NB_AT(dx6, ldd, 6);
#pragma unroll
  for (int k = 0; k < 6; ++k) o[k] = acc ? o[k] + dx[k] : dx[k];
}
__global__ void k_r2aa_fwd(const float* __restrict__ R, int ldr, float* __restrict__ aa, int lda, int per_row, int B,
                           int nz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float r[9], a[3];
  const float* ri = NB_AT(R, ldr, 9);
#pragma unroll
  for (int k = 0; k < 9; ++k) r[k] = ri[k];
  rotmat2aa_fwd(r, a, nz != 0);
  float* o = NB_AT(aa, lda, 3);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
}
__global__ void k_r2aa_bwd(const float* __restrict__ R, int ldr, const float* __restrict__ daa, int lda,
                           float* __restrict__ dR, int ldd, int per_row, int acc, int B, int nz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float r[9], g[3], o[9];
  const float* ri = NB_AT(R, ldr, 9);
  const float* gi = NB_AT(daa, lda, 3);
#pragma unroll
  for (int k = 0; k < 9; ++k) r[k] = ri[k];
  g[0] = gi[0]; g[1] = gi[1]; g[2] = gi[2];
  rotmat2aa_bwd(r, g, o, nz != 0);
  float* p = NB_AT(dR, ldd, 9);
#pragma unroll
  for (int k = 0; k < 9; ++k) p[k] = acc ? p[k] + o[k] : o[k];
}
__global__ void k_rod_fwd(const float* __restrict__ aa, int lda, float* __restrict__ R, int ldr, int per_row, int B,
                          int kind) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float* ai = NB_AT(aa, lda, 3);
  float a[3] = {ai[0], ai[1], ai[2]}, r[9];
  if (kind == 0) rodrigues_quat_fwd(a, r); else rodrigues_skew_fwd(a, r);
  float* ro = NB_AT(R, ldr, 9);
#pragma unroll
  for (int k = 0; k < 9; ++k) ro[k] = r[k];
}
__global__ void k_rod_bwd(const float* __restrict__ aa, int lda, const float* __restrict__ dR, int ldr,
                          float* __restrict__ daa, int ldd, int per_row, int acc, int B, int kind) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float* ai = NB_AT(aa, lda, 3);
  const float* gi = NB_AT(dR, ldr, 9);
  float a[3] = {ai[0], ai[1], ai[2]}, g[9], o[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) g[k] = gi[k];
  if (kind == 0) rodrigues_quat_bwd(a, g, o); else rodrigues_skew_bwd(a, g, o);
  float* p = NB_AT(daa, ldd, 3);
#pragma unroll
  for (int k = 0; k < 3; ++k) p[k] = acc ? p[k] + o[k] : o[k];
}
__global__ void k_pose_feature(const float* __restrict__ R, float* __restrict__ F, int ldf, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * 207) return;
  int n = i / 207, k = i % 207;
  int e = k % 9;
  F[(size_t)n * ldf + k] = R[(size_t)n * 216 + 9 + k] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
}

#define NB_1D(B) nb::cdiv((B), 128), 128, 0, nb::S(stream)
#define NB_ROT_ARGS_OK (B >= 0 && per_row >= 1)

extern "C" int nemo_b200_rot6d_fwd(const float* x6, int ld6, float* R, int ldr, int per_row, int B, void* stream) {
  if (B == 0) return 0;
  NB_REQUIRE(x6 && R && NB_ROT_ARGS_OK, "rot6d_fwd: bad args");
  k_rot6d_fwd<<<NB_1D(B)>>>(x6, ld6, R, ldr, per_row, B);
  NB_LAUNCH_CHECK("rot6d_fwd");
}
extern "C" int nemo_b200_rot6d_bwd(const float* x6, int ld6, const float* dR, int ldr, float* dx6, int ldd,
                                   int per_row, int accumulate, int B, void* stream) {
  if (B == 0) return 0;
  NB_REQUIRE(x6 && dR && dx6 && NB_ROT_ARGS_OK, "rot6d_bwd: bad args");
  k_rot6d_bwd<<<NB_1D(B)>>>(x6, ld6, dR, ldr, dx6, ldd, per_row, accumulate, B);
  NB_LAUNCH_CHECK("rot6d_bwd");
}
extern "C" int nemo_b200_rotmat2aa_fwd(const float* R, int ldr, float* aa, int lda, int per_row, int B,
                                       int nan_to_zero, void* stream) {
  if (B == 0) return 0;
  NB_REQUIRE(R && aa && NB_ROT_ARGS_OK, "rotmat2aa_fwd: bad args");
  k_r2aa_fwd<<<NB_1D(B)>>>(R, ldr, aa, lda, per_row, B, nan_to_zero);
  NB_LAUNCH_CHECK("rotmat2aa_fwd");
}
extern "C" int nemo_b200_rotmat2aa_bwd(const float* R, int ldr, const float* daa, int lda, float* dR, int ldd,
                                       int per_row, int accumulate, int B, int nan_to_zero, void* stream) {
  if (B == 0) return 0;
  NB_REQUIRE(R && daa && dR && NB_ROT_ARGS_OK, "rotmat2aa_bwd: bad args");
  k_r2aa_bwd<<<NB_1D(B)>>>(R, ldr, daa, lda, dR, ldd, per_row, accumulate, B, nan_to_zero);
  NB_LAUNCH_CHECK("rotmat2aa_bwd");
}
extern "C" int nemo_b200_rodrigues_fwd(const float* aa, int lda, float* R, int ldr, int per_row, int B, int kind,
                                       void* stream) {
  if (B == 0) return 0;
  NB_REQUIRE(aa && R && NB_ROT_ARGS_OK && (kind == 0 || kind == 1), "rodrigues_fwd: bad args");
  k_rod_fwd<<<NB_1D(B)>>>(aa, lda, R, ldr, per_row, B, kind);
  NB_LAUNCH_CHECK("rodrigues_fwd");
}
extern "C" int nemo_b200_rodrigues_bwd(const float* aa, int lda, const float* dR, int ldr, float* daa, int ldd,
                                       int per_row, int accumulate, int B, int kind, void* stream) {
  if (B == 0) return 0;
  NB_REQUIRE(aa && dR && daa && NB_ROT_ARGS_OK && (kind == 0 || kind == 1), "rodrigues_bwd: bad args");
  k_rod_bwd<<<NB_1D(B)>>>(aa, lda, dR, ldr, daa, ldd, per_row, accumulate, B, kind);
  NB_LAUNCH_CHECK("rodrigues_bwd");
}
extern "C" int nemo_b200_pose_feature(const float* R, float* F, int ldf, int N, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(R && F && N >= 0 && ldf >= 207, "pose_feature: bad args");
  k_pose_feature<<<NB_1D((long long)N * 207)>>>(R, F, ldf, N);
  NB_LAUNCH_CHECK("pose_feature");
}

// (the kinematic chain lives in chain.cu)

// =============================================================================== joint gather
__global__ void k_jgather_fwd(const float* __restrict__ JP, const float* __restrict__ JX, int nx_raw, int np,
                              const int* __restrict__ jmap, int nout, const float* __restrict__ transl,
                              const float* __restrict__ rowsum, float* __restrict__ out, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * nout * 3) return;
  int c = i % 3, k = (i / 3) % nout, n = i / (3 * nout);
  int r = jmap[k];
  float v, f = 1.f;
  if (r < NJ) v = JP[(size_t)n * 72 + r * 3 + c];
  else {
    v = JX[((size_t)n * nx_raw + (r - NJ)) * 3 + c];
    if (r - NJ >= np && rowsum) f = rowsum[r - NJ - np];
  }
  if (transl) v += f * transl[(size_t)n * 3 + c];
  out[i] = v;
}
__global__ void k_jgather_bwd(const float* __restrict__ dout, int nx_raw, int np, const int* __restrict__ jmap,
                              int nout, const float* __restrict__ rowsum, float* __restrict__ dJP,
                              float* __restrict__ dJX, float* __restrict__ dtr, int N) {
  int nraw = NJ + nx_raw;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * (nraw + 1) * 3) return;
  int c = i % 3, r = (i / 3) % (nraw + 1), n = i / (3 * (nraw + 1));
  const float* d = dout + (size_t)n * nout * 3 + c;
  if (r == nraw) {  // translation
    if (!dtr) return;
    float s = 0.f;
    for (int k = 0; k < nout; ++k) {
      int rr = jmap[k];
      float f = (rr - NJ >= np && rowsum) ? rowsum[rr - NJ - np] : 1.f;
      s += f * d[k * 3];
    }
    dtr[(size_t)n * 3 + c] = s;
    return;
  }
  float s = 0.f;
  for (int k = 0; k < nout; ++k)
    if (jmap[k] == r) s += d[k * 3];
  if (r < NJ) dJP[(size_t)n * 72 + r * 3 + c] = s;
  else dJX[((size_t)n * nx_raw + (r - NJ)) * 3 + c] = s;
}
extern "C" int nemo_b200_joints_gather_fwd(const float* JP, const float* JX, int nraw_x, int np, const int* joint_map,
                                           int nout, const float* transl, const float* rowsum, float* out, int N,
                                           void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(JP && joint_map && out && N >= 0 && nout > 0 && nraw_x >= 0 && np >= 0 && np <= nraw_x &&
                 (nraw_x == 0 || JX), "joints_gather_fwd: bad args");
  k_jgather_fwd<<<NB_1D((long long)N * nout * 3)>>>(JP, JX, nraw_x, np, joint_map, nout, transl, rowsum, out, N);
  NB_LAUNCH_CHECK("joints_gather_fwd");
}
extern "C" int nemo_b200_joints_gather_bwd(const float* dout, int nraw_x, int np, const int* joint_map, int nout,
                                           const float* rowsum, float* dJP, float* dJX, float* dtransl, int N,
                                           void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(dout && joint_map && dJP && N >= 0 && nout > 0 && (nraw_x == 0 || dJX), "joints_gather_bwd: bad args");
  k_jgather_bwd<<<NB_1D((long long)N * (NJ + nraw_x + 1) * 3)>>>(dout, nraw_x, np, joint_map, nout, rowsum, dJP, dJX,
                                                                 dtransl, N);
  NB_LAUNCH_CHECK("joints_gather_bwd");
}

// =============================================================================== projection (API)
__global__ void k_project_fwd(const float* __restrict__ pts, const float* __restrict__ rot,
                              const float* __restrict__ tr, const float* __restrict__ focal,
                              const float* __restrict__ ctr, float eps, float* __restrict__ out, int B, int P) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * P) return;
  int b = i / P;
  float R[9], t[3], p[3], uv[2];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = rot[(size_t)b * 9 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) { t[k] = tr[(size_t)b * 3 + k]; p[k] = pts[(size_t)i * 3 + k]; }
  project_point(R, t, focal[b], ctr[2 * b], ctr[2 * b + 1], eps, p, uv, nullptr);
  out[(size_t)i * 2] = uv[0]; out[(size_t)i * 2 + 1] = uv[1];
}
__global__ void k_project_bwd(const float* __restrict__ pts, const float* __restrict__ rot,
                              const float* __restrict__ tr, const float* __restrict__ focal, float eps,
                              const float* __restrict__ dout, float* __restrict__ dpts, float* __restrict__ drot,
                              float* __restrict__ dtr, int B, int P) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float R[9], t[3], gR[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, gt[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = rot[(size_t)b * 9 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) t[k] = tr[(size_t)b * 3 + k];
  float f = focal[b];
  for (int j = 0; j < P; ++j) {
    size_t i = (size_t)b * P + j;
    float p[3] = {pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]}, uv[2], cam[3], dc[3];
    project_point(R, t, f, 0.f, 0.f, eps, p, uv, cam);
    float duv[2] = {dout[i * 2], dout[i * 2 + 1]};
    // the principal point only enters through eps; its (negligible) term needs cx,cy -> pass 0 when eps == 0
    project_point_bwd(cam, f, 0.f, 0.f, eps, duv, dc);
    if (dpts) {
#pragma unroll
      for (int c = 0; c < 3; ++c) dpts[i * 3 + c] = R[c] * dc[0] + R[3 + c] * dc[1] + R[6 + c] * dc[2];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      gt[r] += dc[r];
#pragma unroll
      for (int c = 0; c < 3; ++c) gR[3 * r + c] += dc[r] * p[c];
    }
  }
  if (drot)
    for (int k = 0; k < 9; ++k) drot[(size_t)b * 9 + k] = gR[k];
  if (dtr)
    for (int k = 0; k < 3; ++k) dtr[(size_t)b * 3 + k] = gt[k];
}
extern "C" int nemo_b200_project_fwd(const float* points, const float* rot, const float* trans, const float* focal,
                                     const float* centre, float eps, float* out, int B, int P, void* stream) {
  if (B == 0) return 0;
  NB_REQUIRE(points && rot && trans && focal && centre && out && B >= 0 && P > 0, "project_fwd: bad args");
  k_project_fwd<<<NB_1D((long long)B * P)>>>(points, rot, trans, focal, centre, eps, out, B, P);
  NB_LAUNCH_CHECK("project_fwd");
}
extern "C" int nemo_b200_project_bwd(const float* points, const float* rot, const float* trans, const float* focal,
                                     float eps, const float* dout, float* dpoints, float* drot, float* dtrans, int B,
                                     int P, void* stream) {
  if (B == 0) return 0;
  NB_REQUIRE(points && rot && trans && focal && dout && B >= 0 && P > 0, "project_bwd: bad args");
  k_project_bwd<<<NB_1D(B)>>>(points, rot, trans, focal, eps, dout, dpoints, drot, dtrans, B, P);
  NB_LAUNCH_CHECK("project_bwd");
}

// =============================================================================== camera + 2D loss (fit path)
struct CamCtx { float R[9], t[3]; };
__device__ __forceinline__ void load_cam(const float* __restrict__ cams, int v, CamCtx& c) {
  float x6[6];
#pragma unroll
  for (int k = 0; k < 3; ++k) c.t[k] = cams[v * 9 + k];
#pragma unroll
  for (int k = 0; k < 6; ++k) x6[k] = cams[v * 9 + 3 + k];
  rot6d_fwd(x6, c.R);
}

__global__ void k_kp_fwd(const float* __restrict__ joints, const int64_t* __restrict__ vi,
                         const int64_t* __restrict__ fi, const float* __restrict__ cams,
                         const float* __restrict__ gt, const float* __restrict__ bbox, float focal, float cx, float cy,
                         float rho, int mode, int reduction, float* __restrict__ p2d,
                         float* __restrict__ loss_all, float* __restrict__ vacc, int N, int P, int V, int F) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int v = (int)vi[n], f = (int)fi[n];
  CamCtx c;
  load_cam(cams, v, c);
  const float* g = gt + ((size_t)v * F + f) * P * 3;
  float size = bbox ? bbox[(size_t)v * F + f] : 1.f;
  int nc = kl_ncomp(mode);
  float acc = 0.f;
  for (int j = 0; j < P; ++j) {
    float p[3] = {joints[((size_t)n * P + j) * 3], joints[((size_t)n * P + j) * 3 + 1], joints[((size_t)n * P + j) * 3 + 2]};
    float uv[2], out[2];
    project_point(c.R, c.t, focal, cx, cy, 0.f, p, uv, nullptr);
    float conf = g[j * 3 + 2];
    keypoint_loss_elem(mode, uv[0] - g[j * 3], uv[1] - g[j * 3 + 1], conf > 0.5f ? 1.f : 0.f, size, rho, out, nullptr,
                       nullptr);
    if (p2d) { p2d[((size_t)n * P + j) * 2] = uv[0]; p2d[((size_t)n * P + j) * 2 + 1] = uv[1]; }
    for (int k = 0; k < nc; ++k) {
      if (loss_all) loss_all[((size_t)n * P + j) * nc + k] = out[k];
      acc += reduction == 0 ? out[k] * conf : out[k];
    }
  }
  atomicAdd(vacc + 2 * v, acc);
  atomicAdd(vacc + 2 * v + 1, 1.f);
}

__global__ void k_kp_bwd(const float* __restrict__ joints, const int64_t* __restrict__ vi,
                         const int64_t* __restrict__ fi, const float* __restrict__ cams,
                         const float* __restrict__ gt, const float* __restrict__ bbox, float focal, float cx, float cy,
                         float rho, int mode, int reduction, const float* __restrict__ vacc, float weight,
                         float* __restrict__ dj, float* __restrict__ dcam, float* __restrict__ loss_out, int N, int P,
                         int V, int F) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int nc = kl_ncomp(mode);
  int nviews = 0;
  float kp = 0.f;
  for (int u = 0; u < V; ++u) {
    float cnt = vacc[2 * u + 1];
    if (cnt > 0.f) { nviews++; kp += reduction == 0 ? vacc[2 * u] / (cnt * P * nc) : vacc[2 * u]; }
  }
  if (reduction != 0) kp /= (float)N * P * nc;
  else kp = nviews ? kp / nviews : 0.f;
  if (n == 0 && loss_out) loss_out[0] = kp;
  if (n >= N) return;
  int v = (int)vi[n], f = (int)fi[n];
  CamCtx c;
  load_cam(cams, v, c);
  const float* g = gt + ((size_t)v * F + f) * P * 3;
  float size = bbox ? bbox[(size_t)v * F + f] : 1.f;
  float coef = reduction == 0 ? weight / (nviews * vacc[2 * v + 1] * P * nc) : weight / ((float)N * P * nc);
  float gR[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, gt3[3] = {0, 0, 0};
  for (int j = 0; j < P; ++j) {
    size_t o = ((size_t)n * P + j) * 3;
    float p[3] = {joints[o], joints[o + 1], joints[o + 2]};
    float uv[2], out[2], cam[3], dr[2], dc[3];
    project_point(c.R, c.t, focal, cx, cy, 0.f, p, uv, cam);
    float conf = g[j * 3 + 2];
    float wc = reduction == 0 ? coef * conf : coef;
    float w[2] = {wc, wc};
    keypoint_loss_elem(mode, uv[0] - g[j * 3], uv[1] - g[j * 3 + 1], conf > 0.5f ? 1.f : 0.f, size, rho, out, w, dr);
    project_point_bwd(cam, focal, cx, cy, 0.f, dr, dc);
#pragma unroll
    for (int k = 0; k < 3; ++k) dj[o + k] = c.R[k] * dc[0] + c.R[3 + k] * dc[1] + c.R[6 + k] * dc[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      gt3[r] += dc[r];
#pragma unroll
      for (int k = 0; k < 3; ++k) gR[3 * r + k] += dc[r] * p[k];
    }
  }
  if (dcam) {
#pragma unroll
    for (int k = 0; k < 3; ++k) atomicAdd(dcam + v * 12 + k, gt3[k]);
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(dcam + v * 12 + 3 + k, gR[k]);
  }
}
__global__ void k_cam_finish(const float* __restrict__ cams, const float* __restrict__ acc, float* __restrict__ dcams,
                             int accumulate, int V) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  float x6[6], dx[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) x6[k] = cams[v * 9 + 3 + k];
  rot6d_bwd(x6, acc + v * 12 + 3, dx);
#pragma unroll
  for (int k = 0; k < 3; ++k) dcams[v * 9 + k] = (accumulate ? dcams[v * 9 + k] : 0.f) + acc[v * 12 + k];
#pragma unroll
  for (int k = 0; k < 6; ++k) dcams[v * 9 + 3 + k] = (accumulate ? dcams[v * 9 + 3 + k] : 0.f) + dx[k];
}
extern "C" int nemo_b200_kp_loss_fwd(const float* joints, const int64_t* view_idx, const int64_t* frame_idx,
                                     const float* cameras, const float* gt, const float* bbox, float focal, float cx,
                                     float cy, float rho, int loss_type, int reduction, float* points2d,
                                     float* loss_all, float* view_acc, int N, int P, int V, int F, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(joints && view_idx && frame_idx && cameras && gt && view_acc && N >= 0 && P > 0 && V > 0 && F > 0 &&
                 loss_type >= 0 && loss_type <= 5, "kp_loss_fwd: bad args");
  NB_REQUIRE(!(loss_type == 1 || loss_type == 5) || bbox, "kp_loss_fwd: resized loss needs bbox");
  k_kp_fwd<<<NB_1D(N)>>>(joints, view_idx, frame_idx, cameras, gt, bbox, focal, cx, cy, rho, loss_type, reduction,
                         points2d, loss_all, view_acc, N, P, V, F);
  NB_LAUNCH_CHECK("kp_loss_fwd");
}
extern "C" int nemo_b200_kp_loss_bwd(const float* joints, const int64_t* view_idx, const int64_t* frame_idx,
                                     const float* cameras, const float* gt, const float* bbox, float focal, float cx,
                                     float cy, float rho, int loss_type, int reduction, const float* view_acc,
                                     float weight, float* djoints, float* dcam_acc, float* loss_out, int N, int P,
                                     int V, int F, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(joints && view_idx && frame_idx && cameras && gt && view_acc && djoints && N >= 0 && P > 0 && V > 0 &&
                 F > 0 && loss_type >= 0 && loss_type <= 5, "kp_loss_bwd: bad args");
  k_kp_bwd<<<NB_1D(N)>>>(joints, view_idx, frame_idx, cameras, gt, bbox, focal, cx, cy, rho, loss_type, reduction,
                         view_acc, weight, djoints, dcam_acc, loss_out, N, P, V, F);
  NB_LAUNCH_CHECK("kp_loss_bwd");
}
extern "C" int nemo_b200_cam_finish(const float* cameras, const float* dcam_acc, float* dcameras, int accumulate,
                                    int V, void* stream) {
  NB_REQUIRE(cameras && dcam_acc && dcameras && V > 0, "cam_finish: bad args");
  k_cam_finish<<<NB_1D(V)>>>(cameras, dcam_acc, dcameras, accumulate, V);
  NB_LAUNCH_CHECK("cam_finish");
}

// =============================================================================== 3D pose loss
__global__ void k_pose3d(const float* __restrict__ pose, int ldp, const int64_t* __restrict__ vi,
                         const int64_t* __restrict__ fi, const float* __restrict__ tgt, const float* __restrict__ mask,
                         float rho, float weight, float* __restrict__ loss_acc, float* __restrict__ dpose, int ldd,
                         int acc, int N, int D, int F) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (i < N * D) {
    int n = i / D, d = i % D;
    size_t vf = (size_t)vi[n] * F + fi[n];
    float r = pose[(size_t)n * ldp + d] - tgt[vf * D + d];
    float m = mask[vf] > 0.5f ? 1.f : 0.f;
    float rho2 = rho * rho, s = r * r;
    l = m * rho2 * (s / (s + rho2)) / ((float)N * D);
    if (dpose) {
      float g = weight / ((float)N * D) * m * rho2 * rho2 * 2.f * r / ((s + rho2) * (s + rho2));
      float* o = dpose + (size_t)n * ldd + d;
      *o = acc ? *o + g : g;
    }
  }
  l = warp_sum(l);
  __shared__ float red[4];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x == 0 && loss_acc) atomicAdd(loss_acc, red[0] + red[1] + red[2] + red[3]);
}
extern "C" int nemo_b200_pose3d_loss(const float* pose, int ldp, const int64_t* view_idx, const int64_t* frame_idx,
                                     const float* target, const float* mask, float rho, float weight, float* loss_acc,
                                     float* dpose, int ldd, int accumulate, int N, int D, int F, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(pose && view_idx && frame_idx && target && mask && N >= 0 && D > 0 && F > 0 && ldp >= D,
             "pose3d_loss: bad args");
  k_pose3d<<<NB_1D((long long)N * D)>>>(pose, ldp, view_idx, frame_idx, target, mask, rho, weight, loss_acc, dpose,
                                        ldd, accumulate, N, D, F);
  NB_LAUNCH_CHECK("pose3d_loss");
}

// =============================================================================== K1: phase warp + RBF + code
// y(u) = mean_k sigmoid(relu(scale_k) (u - relu(shift_k)))
__device__ __forceinline__ float mono_y_lane(const float* __restrict__ sh, const float* __restrict__ sc, int K,
                                             float u, int lane) {
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += sigmoidf_(fmaxf(sc[k], 0.f) * (u - fmaxf(sh[k], 0.f)));
  return warp_sum(s) / (float)K;
}
// y(0), y(1) of every view's phase warp: one block per view (a warp per view walked its K nodes in K/32 dependent
// rounds: 7 us of latency at the head of every step)
__global__ void __launch_bounds__(128) k_mono_norm(const float* __restrict__ shifts, const float* __restrict__ scales,
                                                   int K, float* __restrict__ yn, int V) {
  __shared__ float red[2][4];
  const int v = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float* sh = shifts + (size_t)v * K;
  const float* sc = scales + (size_t)v * K;
  float s0 = 0.f, s1 = 0.f;
  for (int k = threadIdx.x; k < K; k += 128) {
    const float a = fmaxf(sc[k], 0.f), b = fmaxf(sh[k], 0.f);
    s0 += sigmoidf_(a * (0.f - b));
    s1 += sigmoidf_(a * (1.f - b));
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if (lane == 0) { red[0][w] = s0; red[1][w] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float y0 = ((red[0][0] + red[0][1]) + (red[0][2] + red[0][3])) / (float)K;
    const float y1 = ((red[1][0] + red[1][1]) + (red[1][2] + red[1][3])) / (float)K;
    yn[3 * v] = y0; yn[3 * v + 1] = y1; yn[3 * v + 2] = y1 - y0 + 1e-6f;
  }
}
// one warp per sample; sample N (the extra row) is the phase-0 / zero-code input of trans_0 (:3755-3766)
__global__ void k_embed_fwd(const float* __restrict__ raw, const int64_t* __restrict__ vi,
                            const float* __restrict__ shifts, const float* __restrict__ scales, int K,
                            const float* __restrict__ centres, const float* __restrict__ ls, int D, int basis,
                            const float* __restrict__ codes, int C, const float* __restrict__ noise,
                            float* __restrict__ X, int ldx, float* __restrict__ warped, const float* __restrict__ yn,
                            int N) {
  int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n > N) return;
  float ph = 0.f;
  int v = 0;
  if (n < N) {
    v = (int)vi[n];
    float y = mono_y_lane(shifts + (size_t)v * K, scales + (size_t)v * K, K, raw[n], lane);
    ph = (y - yn[3 * v]) / yn[3 * v + 2];
    if (lane == 0) warped[n] = ph;
  }
  float* x = X + (size_t)n * ldx;
  int nf = D > 0 ? D : 1;
  if (D > 0) {
    for (int i = lane; i < D; i += 32) {
      float d = ph - centres[i];
      x[i] = rbf_basis(basis, d * d / expf(ls[i]), nullptr);
    }
  } else if (lane == 0) x[0] = ph;
  for (int c = lane; c < C; c += 32)
    x[nf + c] = (n < N) ? codes[(size_t)v * C + c] + (noise ? noise[(size_t)n * C + c] : 0.f) : 0.f;
}
// block = 8 warps, each warp takes a run of the block's samples; dlog_sigmas reduced in the block first
constexpr int EB_KPL = 8;   // phase-network nodes per lane held in registers: K <= 256
__global__ void __launch_bounds__(256) k_embed_bwd(const float* __restrict__ raw, const int64_t* __restrict__ vi,
                                                   const float* __restrict__ shifts, const float* __restrict__ scales,
                                                   int K, const float* __restrict__ centres,
                                                   const float* __restrict__ ls, int D, int basis, int C,
                                                   const float* __restrict__ warped, const float* __restrict__ yn,
                                                   const float* __restrict__ dX, int ldx, float* __restrict__ dls,
                                                   float* __restrict__ dsh, float* __restrict__ dsc,
                                                   float* __restrict__ dcodes, float* __restrict__ gy01, int N,
                                                   int per_block) {
  extern __shared__ float sm[];  // [8][D]
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int nf = D > 0 ? D : 1;
  float* myls = sm + (size_t)w * (D > 0 ? D : 1);
  for (int i = lane; i < D; i += 32) myls[i] = 0.f;
  int n0 = blockIdx.x * per_block;
  float asc[EB_KPL], ash[EB_KPL], g0acc = 0.f, g1acc = 0.f, cacc = 0.f;
#pragma unroll
  for (int t = 0; t < EB_KPL; ++t) asc[t] = ash[t] = 0.f;
  int cur_v = -1;
  auto flush = [&]() {
    if (cur_v < 0) return;
#pragma unroll
    for (int t = 0; t < EB_KPL; ++t) {
      const int k = lane + 32 * t;
      if (k < K) {
        if (asc[t] != 0.f) atomicAdd(dsc + (size_t)cur_v * K + k, asc[t]);
        if (ash[t] != 0.f) atomicAdd(dsh + (size_t)cur_v * K + k, ash[t]);
      }
      asc[t] = ash[t] = 0.f;
    }
    if (lane == 0) {   // every lane carries the same sums
      atomicAdd(gy01 + 2 * cur_v, g0acc);
      atomicAdd(gy01 + 2 * cur_v + 1, g1acc);
    }
    if (dcodes && lane < C && cacc != 0.f) atomicAdd(dcodes + (size_t)cur_v * C + lane, cacc);
    g0acc = g1acc = cacc = 0.f;
  };
  // a warp takes a contiguous run of the block's samples
  const int per_warp = per_block / 8;
  for (int s = w * per_warp; s < (w + 1) * per_warp; ++s) {
    int n = n0 + s;
    if (n > N) break;
    const float* dx = dX + (size_t)n * ldx;
    float ph = n < N ? warped[n] : 0.f;
    float gph = 0.f;
    if (D > 0) {
      for (int i = lane; i < D; i += 32) {
        float d = ph - centres[i];
        float e = expf(-ls[i]);
        float a = d * d * e, da;
        rbf_basis(basis, a, &da);
        float g = dx[i] * da;
        myls[i] += -a * g;
        gph += g * 2.f * d * e;
      }
      gph = warp_sum(gph);
    } else gph = dx[0];
    if (n >= N) continue;
    int v = (int)vi[n];
    const float dcode_n = (dcodes && lane < C) ? dx[nf + lane] : 0.f;   // instance codes: C <= 32 (checked by the host)
    // monotonic network backward.  The per-node gradients of the sample's view accumulate in registers while the warp's
    // consecutive samples share a view (batches are usually sorted by view or cover whole views) and go out as one
    // atomic per node when the view changes: 4096 samples x 200 nodes x 2 atomics on 3200 addresses were the cost of
    // this kernel.
    float den = yn[3 * v + 2];
    float gy = gph / den;
    if (v != cur_v) {
      flush();
      cur_v = v;
    }
    g0acc += gph * (ph - 1.f) / den;
    g1acc += -gph * ph / den;
    cacc += dcode_n;
    const float* sh = shifts + (size_t)v * K;
    const float* sc = scales + (size_t)v * K;
    float u = raw[n];
#pragma unroll
    for (int t = 0; t < EB_KPL; ++t) {
      const int k = lane + 32 * t;
      if (k < K) {
        float a = fmaxf(sc[k], 0.f), b = fmaxf(sh[k], 0.f);
        float sg = sigmoidf_(a * (u - b));
        float dz = gy * sg * (1.f - sg) / (float)K;
        if (sc[k] > 0.f) asc[t] += dz * (u - b);
        if (sh[k] > 0.f) ash[t] += -dz * a;
      }
    }
  }
  flush();
  __syncthreads();
  if (D > 0 && dls)
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      float s = 0.f;
      for (int q = 0; q < 8; ++q) s += sm[(size_t)q * D + i];
      atomicAdd(dls + i, s);
    }
}
// adds the y(0), y(1) normaliser terms of the phase warp: one block per view, one node per thread and round
__global__ void __launch_bounds__(128) k_mono_finish(const float* __restrict__ shifts, const float* __restrict__ scales,
                                                     int K, const float* __restrict__ gy01, float* __restrict__ dsh,
                                                     float* __restrict__ dsc, int V) {
  const int v = blockIdx.x;
  float g0 = gy01[2 * v], g1 = gy01[2 * v + 1];
  for (int k = threadIdx.x; k < K; k += 128) {
    float scv = scales[(size_t)v * K + k], shv = shifts[(size_t)v * K + k];
    float a = fmaxf(scv, 0.f), b = fmaxf(shv, 0.f);
    float s0 = sigmoidf_(a * (0.f - b)), s1 = sigmoidf_(a * (1.f - b));
    float d0 = g0 * s0 * (1.f - s0) / (float)K, d1 = g1 * s1 * (1.f - s1) / (float)K;
    if (scv > 0.f) dsc[(size_t)v * K + k] += d0 * (0.f - b) + d1 * (1.f - b);
    if (shv > 0.f) dsh[(size_t)v * K + k] += -(d0 + d1) * a;
  }
}
extern "C" int nemo_b200_embed_fwd(const float* raw_phase, const int64_t* view_idx, const float* shifts,
                                   const float* scales, int K, const float* centres, const float* log_sigmas, int D,
                                   int basis, const float* codes, int C, const float* noise, float* X, int ldx,
                                   float* warped, float* yn, int N, int V, void* stream) {
  NB_REQUIRE(raw_phase && view_idx && shifts && scales && X && warped && yn && K > 0 && N >= 0 && V > 0 && D >= 0 &&
                 C >= 0 && ldx >= (D > 0 ? D : 1) + C && (D == 0 || (centres && log_sigmas)) && (C == 0 || codes) &&
                 basis >= 0 && basis <= 10, "embed_fwd: bad args");
  k_mono_norm<<<V, 128, 0, S(stream)>>>(shifts, scales, K, yn, V);
  NB_LAUNCH_CHECK_CONT("mono_norm");
  k_embed_fwd<<<cdiv(N + 1, 4), 128, 0, S(stream)>>>(raw_phase, view_idx, shifts, scales, K, centres, log_sigmas, D,
                                                     basis, codes, C, noise, X, ldx, warped, yn, N);
  NB_LAUNCH_CHECK("embed_fwd");
}
extern "C" int nemo_b200_embed_bwd(const float* raw_phase, const int64_t* view_idx, const float* shifts,
                                   const float* scales, int K, const float* centres, const float* log_sigmas, int D,
                                   int basis, int C, const float* warped, const float* yn, const float* dX, int ldx,
                                   float* dlog_sigmas, float* dshifts, float* dscales, float* dcodes, float* scratch,
                                   int N, int V, void* stream) {
  NB_REQUIRE(raw_phase && view_idx && shifts && scales && warped && yn && dX && dshifts && dscales && scratch &&
                 K > 0 && K <= 32 * EB_KPL && C <= 32 && N >= 0 && V > 0 && D >= 0 &&
                 (D == 0 || (centres && log_sigmas && dlog_sigmas)),
             "embed_bwd: bad args (at most 256 phase-network nodes, 32 instance-code entries)");
  NB_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * V, S(stream)));
  // samples per block (8 warps, a contiguous run each): 64 left 65 blocks for 148 SMs and eight dependent samples per
  // warp -- measured at N = 4096: 0.046 ms (64), 0.033 (32), 0.029 (16).  NEMO_B200_EMBED_PB: measurement switch
  static const int per_block = [] {
    const char* e = getenv("NEMO_B200_EMBED_PB");
    const int v = e ? atoi(e) : 16;
    return (v == 8 || v == 16 || v == 32 || v == 64) ? v : 16;
  }();
  size_t smem = sizeof(float) * 8 * (D > 0 ? D : 1);
  k_embed_bwd<<<cdiv(N + 1, per_block), 256, smem, S(stream)>>>(raw_phase, view_idx, shifts, scales, K, centres,
                                                                log_sigmas, D, basis, C, warped, yn, dX, ldx,
                                                                dlog_sigmas, dshifts, dscales, dcodes, scratch, N,
                                                                per_block);
  NB_LAUNCH_CHECK_CONT("embed_bwd");
  k_mono_finish<<<V, 128, 0, S(stream)>>>(shifts, scales, K, scratch, dshifts, dscales, V);
  NB_LAUNCH_CHECK("mono_finish");
}

// ---- standalone layers (module API): nemo.rbf.RBF, monotonic_network.MonotonicNetwork
__global__ void k_rbf_fwd(const float* __restrict__ x, const float* __restrict__ c, const float* __restrict__ ls, int D,
                          int basis, float* __restrict__ out, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * D) return;
  int n = i / D, k = i % D;
  float d = x[n] - c[k];
  out[i] = rbf_basis(basis, d * d / expf(ls[k]), nullptr);
}
__global__ void k_rbf_bwd(const float* __restrict__ x, const float* __restrict__ c, const float* __restrict__ ls, int D,
                          int basis, const float* __restrict__ dout, float* __restrict__ dx, float* __restrict__ dls,
                          int N) {
  int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float g = 0.f;
  for (int k = lane; k < D; k += 32) {
    float d = x[n] - c[k], e = expf(-ls[k]);
    float a = d * d * e, da;
    rbf_basis(basis, a, &da);
    float go = dout[(size_t)n * D + k] * da;
    if (dls) atomicAdd(dls + k, -a * go);
    g += go * 2.f * d * e;
  }
  g = warp_sum(g);
  if (lane == 0 && dx) dx[n] = g;
}
__global__ void k_mono_fwd(const float* __restrict__ x, const float* __restrict__ sh, const float* __restrict__ sc,
                           int K, float* __restrict__ out, int N) {
  int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float y = mono_y_lane(sh, sc, K, x[n], lane), y0 = mono_y_lane(sh, sc, K, 0.f, lane),
        y1 = mono_y_lane(sh, sc, K, 1.f, lane);
  if (lane == 0) out[n] = (y - y0) / (y1 - y0 + 1e-6f);
}
__global__ void k_mono_bwd(const float* __restrict__ x, const float* __restrict__ sh, const float* __restrict__ sc,
                           int K, const float* __restrict__ dout, float* __restrict__ dx, float* __restrict__ dsh,
                           float* __restrict__ dsc, float* __restrict__ gy01, int N) {
  int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float u = x[n];
  float y = mono_y_lane(sh, sc, K, u, lane), y0 = mono_y_lane(sh, sc, K, 0.f, lane),
        y1 = mono_y_lane(sh, sc, K, 1.f, lane);
  float den = y1 - y0 + 1e-6f, out = (y - y0) / den, g = dout[n];
  float gy = g / den, gx = 0.f;
  if (lane == 0) { atomicAdd(gy01, g * (out - 1.f) / den); atomicAdd(gy01 + 1, -g * out / den); }
  for (int k = lane; k < K; k += 32) {
    float a = fmaxf(sc[k], 0.f), b = fmaxf(sh[k], 0.f);
    float sg = sigmoidf_(a * (u - b));
    float dz = gy * sg * (1.f - sg) / (float)K;
    gx += dz * a;
    if (sc[k] > 0.f) atomicAdd(dsc + k, dz * (u - b));
    if (sh[k] > 0.f) atomicAdd(dsh + k, -dz * a);
  }
  gx = warp_sum(gx);
  if (lane == 0 && dx) dx[n] = gx;
}
extern "C" int nemo_b200_rbf_fwd(const float* x, const float* centres, const float* log_sigmas, int D, int basis,
                                 float* out, int N, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(x && centres && log_sigmas && out && D > 0 && N >= 0 && basis >= 0 && basis <= 10, "rbf_fwd: bad args");
  k_rbf_fwd<<<NB_1D((long long)N * D)>>>(x, centres, log_sigmas, D, basis, out, N);
  NB_LAUNCH_CHECK("rbf_fwd");
}
extern "C" int nemo_b200_rbf_bwd(const float* x, const float* centres, const float* log_sigmas, int D, int basis,
                                 const float* dout, float* dx, float* dlog_sigmas, int N, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(x && centres && log_sigmas && dout && D > 0 && N >= 0 && basis >= 0 && basis <= 10, "rbf_bwd: bad args");
  k_rbf_bwd<<<cdiv(N, 4), 128, 0, S(stream)>>>(x, centres, log_sigmas, D, basis, dout, dx, dlog_sigmas, N);
  NB_LAUNCH_CHECK("rbf_bwd");
}
extern "C" int nemo_b200_monotonic_fwd(const float* x, const float* shifts, const float* scales, int K, float* out,
                                       int N, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(x && shifts && scales && out && K > 0 && N >= 0, "monotonic_fwd: bad args");
  k_mono_fwd<<<cdiv(N, 4), 128, 0, S(stream)>>>(x, shifts, scales, K, out, N);
  NB_LAUNCH_CHECK("monotonic_fwd");
}
extern "C" int nemo_b200_monotonic_bwd(const float* x, const float* shifts, const float* scales, int K,
                                       const float* dout, float* dx, float* dshifts, float* dscales, float* scratch,
                                       int N, void* stream) {
  NB_REQUIRE(x && shifts && scales && dout && dshifts && dscales && scratch && K > 0 && N >= 0,
             "monotonic_bwd: bad args");
  NB_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * 2, S(stream)));
  if (N > 0) {
    k_mono_bwd<<<cdiv(N, 4), 128, 0, S(stream)>>>(x, shifts, scales, K, dout, dx, dshifts, dscales, scratch, N);
    NB_LAUNCH_CHECK_CONT("monotonic_bwd");
  }
  k_mono_finish<<<1, 128, 0, S(stream)>>>(shifts, scales, K, scratch, dshifts, dscales, 1);
  NB_LAUNCH_CHECK("mono_finish");
}

// =============================================================================== GMM prior (K8)
// block = 32 samples (lanes) x 8 warps (row chunks); loops over the M components with Psym_m in smem.
constexpr int GD = 69, GDP = 72;
__global__ void __launch_bounds__(256) k_gmm(const float* __restrict__ pose, int ldp, const float* __restrict__ means,
                                             const float* __restrict__ prec, const float* __restrict__ nllw, int M,
                                             float* __restrict__ ll, float gscale, float* __restrict__ dpose, int ldd,
                                             int acc, int N, const float* __restrict__ psym) {
  __shared__ __align__(16) float sP[GD * GDP];     // symmetrised precision of the current component, rows padded to 72
  __shared__ float sD[GDP][33];      // d[j][sample]
  __shared__ float sQ[8][32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int n = blockIdx.x * 32 + lane;
  bool live = n < N;
  for (int i = threadIdx.x; i < GD * (GDP - GD); i += 256) sP[(i / (GDP - GD)) * GDP + GD + i % (GDP - GD)] = 0.f;
  float x[GD];
#pragma unroll
  for (int j = 0; j < GD; ++j) x[j] = live ? pose[(size_t)n * ldp + j] : 0.f;
  float best = 3.4e38f;
  float ybest[9];
#pragma unroll
  for (int r = 0; r < 9; ++r) ybest[r] = 0.f;
  for (int m = 0; m < M; ++m) {
    __syncthreads();
    if (psym) {   // symmetrised rows padded to 72 floats, prepared once (nemo_b200_gmm_prepare): straight 16-byte copies
      const float4* src = reinterpret_cast<const float4*>(psym + (size_t)m * GD * GDP);
      float4* dst = reinterpret_cast<float4*>(sP);
      for (int i = threadIdx.x; i < GD * GDP / 4; i += 256) dst[i] = __ldg(src + i);
    } else {
      for (int i = threadIdx.x; i < GD * GD; i += 256) {
        int a = i / GD, b = i % GD;
        sP[a * GDP + b] = 0.5f * (prec[((size_t)m * GD + a) * GD + b] + prec[((size_t)m * GD + b) * GD + a]);
      }
    }
    float d[GDP];
#pragma unroll
    for (int j = 0; j < GD; ++j) d[j] = x[j] - means[m * GD + j];
#pragma unroll
    for (int j = GD; j < GDP; ++j) d[j] = 0.f;
    if (w == 0) {
#pragma unroll
      for (int j = 0; j < GD; ++j) sD[j][lane] = d[j];
    }
    __syncthreads();
    float y[9], q = 0.f;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      int i = w * 9 + r;
      float a = 0.f;
      if (i < GD) {
        // the row is a warp-wide broadcast (16-byte loads); four independent partial sums: one 72-long dependent FMA
        // chain per row was the critical path of this kernel
        const float4* row = reinterpret_cast<const float4*>(sP + i * GDP);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int j = 0; j < GDP / 4; ++j) {
          const float4 p4 = row[j];
          a0 = fmaf(p4.x, d[4 * j], a0);
          a1 = fmaf(p4.y, d[4 * j + 1], a1);
          a2 = fmaf(p4.z, d[4 * j + 2], a2);
          a3 = fmaf(p4.w, d[4 * j + 3], a3);
        }
        a = (a0 + a1) + (a2 + a3);
        q += a * sD[i][lane];
      }
      y[r] = a;
    }
    sQ[w][lane] = q;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += sQ[k][lane];
    float l = 0.5f * tot - logf(nllw[m]);
    if (l < best) {
      best = l;
#pragma unroll
      for (int r = 0; r < 9; ++r) ybest[r] = y[r];
    }
  }
  if (!live) return;
  if (w == 0 && ll) ll[n] = best;
  if (dpose) {
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      int i = w * 9 + r;
      if (i < GD) {
        float* o = dpose + (size_t)n * ldd + i;
        float g = gscale * ybest[r];
        *o = acc ? *o + g : g;
      }
    }
  }
}
extern "C" int nemo_b200_gmm_prior(const float* pose, int ldp, const float* means, const float* precisions,
                                   const float* nll_weights, int M, int D, float* ll, float gscale, float* dpose,
                                   int ldd, int accumulate, int N, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(pose && means && precisions && nll_weights && M > 0 && D == GD && ldp >= D && N >= 0,
             "gmm_prior: bad args (D must be 69)");
  k_gmm<<<cdiv(N, 32), 256, 0, S(stream)>>>(pose, ldp, means, precisions, nll_weights, M, ll, gscale, dpose, ldd,
                                            accumulate, N, nullptr);
  NB_LAUNCH_CHECK("gmm_prior");
}
// The precision matrices are frozen buffers (hmr/smplify/prior.py:142-160): symmetrise and pad them once.
__global__ void k_gmm_prepare(const float* __restrict__ prec, float* __restrict__ psym, int M) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * GD * GDP) return;
  int m = i / (GD * GDP), a = (i / GDP) % GD, b = i % GDP;
  psym[i] = b < GD ? 0.5f * (prec[((size_t)m * GD + a) * GD + b] + prec[((size_t)m * GD + b) * GD + a]) : 0.f;
}
extern "C" int nemo_b200_gmm_prepare(const float* precisions, int M, int D, float* prec_sym, void* stream) {
  NB_REQUIRE(precisions && prec_sym && M > 0 && D == GD, "gmm_prepare: bad args (D must be 69)");
  k_gmm_prepare<<<NB_1D((long long)M * GD * GDP)>>>(precisions, prec_sym, M);
  NB_LAUNCH_CHECK("gmm_prepare");
}
extern "C" int nemo_b200_gmm_prior_sym(const float* pose, int ldp, const float* means, const float* prec_sym,
                                       const float* nll_weights, int M, int D, float* ll, float gscale, float* dpose,
                                       int ldd, int accumulate, int N, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(pose && means && prec_sym && nll_weights && M > 0 && D == GD && ldp >= D && N >= 0 &&
                 ((uintptr_t)prec_sym & 15) == 0, "gmm_prior_sym: bad args (D must be 69, prec_sym 16-byte aligned)");
  k_gmm<<<cdiv(N, 32), 256, 0, S(stream)>>>(pose, ldp, means, nullptr, nll_weights, M, ll, gscale, dpose, ldd, accumulate,
                                            N, prec_sym);
  NB_LAUNCH_CHECK("gmm_prior_sym");
}

// =============================================================================== VPoser KL head
__global__ void k_vp_kl(const float* __restrict__ head, int L, float* __restrict__ sigma, float* __restrict__ kl_acc,
                        float gscale, float* __restrict__ dhead, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float kl = 0.f;
  if (i < N * L) {
    int n = i / L, d = i % L;
    float mu = head[(size_t)n * 2 * L + d], h = head[(size_t)n * 2 * L + L + d];
    float s = softplusf_(h);
    kl = -logf(s) + 0.5f * (s * s + mu * mu) - 0.5f;
    if (sigma) sigma[i] = s;
    if (dhead) {
      dhead[(size_t)n * 2 * L + d] = gscale * mu;
      float dsdh = h > 20.f ? 1.f : sigmoidf_(h);
      dhead[(size_t)n * 2 * L + L + d] = gscale * (s - 1.f / s) * dsdh;
    }
  }
  kl = warp_sum(kl);
  __shared__ float red[4];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = kl;
  __syncthreads();
  if (threadIdx.x == 0 && kl_acc) atomicAdd(kl_acc, red[0] + red[1] + red[2] + red[3]);
}
extern "C" int nemo_b200_vposer_kl(const float* head, int L, float* sigma, float* kl_acc, float gscale, float* dhead,
                                   int N, void* stream) {
  if (N == 0) return 0;
  NB_REQUIRE(head && L > 0 && N >= 0, "vposer_kl: bad args");
  k_vp_kl<<<NB_1D((long long)N * L)>>>(head, L, sigma, kl_acc, gscale, dhead, N);
  NB_LAUNCH_CHECK("vposer_kl");
}

// =============================================================================== Adam + utilities
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps, float wd, int dec,
                       float bc1, float bc2s, const int* __restrict__ step_dev) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (step_dev) {  // step count lives on the device so a captured CUDA graph can be replayed
    float t = (float)(*step_dev);
    bc1 = 1.f - powf(b1, t);
    bc2s = sqrtf(1.f - powf(b2, t));
  }
  adam_elem(p + i, g[i], m + i, v + i, lr, b1, b2, eps, wd, dec, bc1, bc2s);
}
__global__ void k_increment(int* c) { *c += 1; }
extern "C" int nemo_b200_adam(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int decoupled, int step,
                              const int* step_dev, void* stream) {
  if (n == 0) return 0;
  NB_REQUIRE(p && g && m && v && (step >= 1 || step_dev), "adam: bad args");
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  k_adam<<<cdiv((long long)n, 256), 256, 0, S(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, decoupled,
                                                         bc1, bc2s, step_dev);
  NB_LAUNCH_CHECK("adam");
}
extern "C" int nemo_b200_increment(int* counter, void* stream) {
  NB_REQUIRE(counter, "increment: bad args");
  k_increment<<<1, 1, 0, S(stream)>>>(counter);
  NB_LAUNCH_CHECK("increment");
}

// One launch for every optimiser of a fit model (the reference steps 3-4 torch.optim.Adam instances, each a loop over
// its parameter tensors: nemo/neural_motion_model.py:3586-3588).  A segment is a contiguous range of one flat parameter
// buffer with its own hyper-parameters and a device-resident count of COMPLETED steps; segments may share a counter.
// Every block reads `*step + 1` as the current step; the counters are advanced by the last block to finish (ticket),
// so no block can see the incremented value and no separate increment kernel is needed.  Segments flagged SKIP are
// left untouched and their counter is not advanced -- torch.optim.Adam skips parameters whose grad is None
// (warm-up leaves linear_out.* without gradient, :3455-3509).
struct AdamTable {
  nemo_b200_adam_seg seg[NEMO_B200_ADAM_MAX_SEGS];
  int first_block[NEMO_B200_ADAM_MAX_SEGS + 1];
  int nseg;
};
__global__ void k_adam_multi(const __grid_constant__ AdamTable t, int* __restrict__ ticket) {
  int s = 0;
  while (s + 1 < t.nseg && (int)blockIdx.x >= t.first_block[s + 1]) ++s;
  const nemo_b200_adam_seg& sg = t.seg[s];
  const size_t i0 = ((size_t)(blockIdx.x - t.first_block[s]) * blockDim.x + threadIdx.x) * 4;
  if (!(sg.flags & NEMO_B200_ADAM_SKIP) && i0 < sg.n) {
    const float step = (float)(*sg.step + 1);
    const float bc1 = 1.f - powf(sg.beta1, step), bc2s = sqrtf(1.f - powf(sg.beta2, step));
    if (i0 + 4 <= sg.n) {  // segments start on 16-byte boundaries (ParamGroup rounds every tensor up to 4 floats)
      float4 p = *reinterpret_cast<float4*>(sg.p + i0), m = *reinterpret_cast<float4*>(sg.m + i0),
             v = *reinterpret_cast<float4*>(sg.v + i0);
      const float4 g = *reinterpret_cast<const float4*>(sg.g + i0);
      adam_elem(&p.x, g.x, &m.x, &v.x, sg.lr, sg.beta1, sg.beta2, sg.eps, sg.weight_decay, sg.decoupled, bc1, bc2s);
      adam_elem(&p.y, g.y, &m.y, &v.y, sg.lr, sg.beta1, sg.beta2, sg.eps, sg.weight_decay, sg.decoupled, bc1, bc2s);
      adam_elem(&p.z, g.z, &m.z, &v.z, sg.lr, sg.beta1, sg.beta2, sg.eps, sg.weight_decay, sg.decoupled, bc1, bc2s);
      adam_elem(&p.w, g.w, &m.w, &v.w, sg.lr, sg.beta1, sg.beta2, sg.eps, sg.weight_decay, sg.decoupled, bc1, bc2s);
      *reinterpret_cast<float4*>(sg.p + i0) = p;
      *reinterpret_cast<float4*>(sg.m + i0) = m;
      *reinterpret_cast<float4*>(sg.v + i0) = v;
    } else {
      for (size_t i = i0; i < sg.n; ++i)
        adam_elem(sg.p + i, sg.g[i], sg.m + i, sg.v + i, sg.lr, sg.beta1, sg.beta2, sg.eps, sg.weight_decay,
                  sg.decoupled, bc1, bc2s);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
      for (int k = 0; k < t.nseg; ++k)
        if ((t.seg[k].flags & NEMO_B200_ADAM_OWNS_STEP) && !(t.seg[k].flags & NEMO_B200_ADAM_SKIP)) *t.seg[k].step += 1;
      *ticket = 0;
    }
  }
}
extern "C" int nemo_b200_adam_multi(const nemo_b200_adam_seg* segs, int nseg, int* ticket, void* stream) {
  NB_REQUIRE(segs && ticket && nseg > 0 && nseg <= NEMO_B200_ADAM_MAX_SEGS, "adam_multi: 1..%d segments",
             NEMO_B200_ADAM_MAX_SEGS);
  AdamTable t;
  t.nseg = nseg;
  int blocks = 0;
  for (int s = 0; s < nseg; ++s) {
    NB_REQUIRE(segs[s].p && segs[s].g && segs[s].m && segs[s].v && segs[s].step, "adam_multi: null buffer in segment %d", s);
    NB_REQUIRE(((uintptr_t)segs[s].p % 16 == 0) && ((uintptr_t)segs[s].g % 16 == 0) && ((uintptr_t)segs[s].m % 16 == 0) &&
               ((uintptr_t)segs[s].v % 16 == 0), "adam_multi: segment %d is not 16-byte aligned", s);
    t.seg[s] = segs[s];
    t.first_block[s] = blocks;
    blocks += max(1, cdiv((long long)segs[s].n, 1024));
  }
  t.first_block[nseg] = blocks;
  k_adam_multi<<<blocks, 256, 0, S(stream)>>>(t, ticket);
  NB_LAUNCH_CHECK("adam_multi");
}
__global__ void k_scale_add(float* __restrict__ y, const float* __restrict__ x, float a, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += a * x[i];
}
extern "C" int nemo_b200_scale_add(float* y, const float* x, float a, size_t n, void* stream) {
  if (n == 0) return 0;
  NB_REQUIRE(y && x, "scale_add: bad args");
  k_scale_add<<<cdiv((long long)n, 256), 256, 0, S(stream)>>>(y, x, a, n);
  NB_LAUNCH_CHECK("scale_add");
}
__global__ void k_act_bwd(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ out, size_t n,
                          float slope) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dy[i] * (y[i] > 0.f ? 1.f : slope);
}
extern "C" int nemo_b200_act_bwd(const float* dy, const float* y, float* out, size_t n, int act, void* stream) {
  if (n == 0) return 0;
  NB_REQUIRE(dy && y && out && (act == 1 || act == 2), "act_bwd: bad args");
  k_act_bwd<<<cdiv((long long)n, 256), 256, 0, S(stream)>>>(dy, y, out, n, act == 1 ? 0.f : 0.01f);
  NB_LAUNCH_CHECK("act_bwd");
}
// see nemo_b200_dy_tail in the header.  Each CTA fills the tail columns of a slice of rows and adds its column sums of
// dtrans to scratch[0..2]; the last CTA to finish (ticket in scratch[3]) writes the phase-0 row.
__global__ void __launch_bounds__(256) k_dy_tail(const float* __restrict__ dtr, float* __restrict__ dY, int ldy, int N,
                                                 float* __restrict__ scratch) {
  __shared__ float red[3][8];
  __shared__ int last;
  float s[3] = {0.f, 0.f, 0.f};
  for (int n = blockIdx.x * 256 + threadIdx.x; n < N; n += gridDim.x * 256) {
    float* row = dY + (size_t)n * ldy;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = dtr ? dtr[(size_t)n * 3 + c] : 0.f;
      row[144 + c] = v;
      s[c] += v;
    }
    for (int c = 147; c < ldy; ++c) row[c] = 0.f;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = warp_sum(s[c]);
    if ((threadIdx.x & 31) == 0) red[c][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    atomicAdd(scratch + threadIdx.x, v);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(reinterpret_cast<int*>(scratch + 3), 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  float* lastrow = dY + (size_t)N * ldy;
  for (int c = threadIdx.x; c < ldy; c += 256)
    lastrow[c] = (c >= 144 && c < 147) ? -__ldcg(scratch + (c - 144)) : 0.f;
}
extern "C" int nemo_b200_dy_tail(const float* dtrans, float* dY, int ldy, int N, float* scratch, void* stream) {
  NB_REQUIRE(dY && scratch && ldy >= 147 && N >= 0, "dy_tail: bad args");
  k_dy_tail<<<max(1, min(cdiv(N, 256), 64)), 256, 0, S(stream)>>>(dtrans, dY, ldy, N, scratch);
  NB_LAUNCH_CHECK("dy_tail");
}
__global__ void k_combine(const float* __restrict__ t, const float* __restrict__ w, int n, float* __restrict__ out) {
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += w[i] * t[i];
  out[0] = s;
}
extern "C" int nemo_b200_combine_losses(const float* terms, const float* weights, int n, float* out, void* stream) {
  NB_REQUIRE(terms && weights && out && n > 0, "combine_losses: bad args");
  k_combine<<<1, 1, 0, S(stream)>>>(terms, weights, n, out);
  NB_LAUNCH_CHECK("combine_losses");
}
