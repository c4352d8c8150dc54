// CPU build of the host+device math headers (rotmath.cuh, pathmath.cuh) so the "-m 'not gpu'" tests can pin
// the exact per-element arithmetic the kernels execute against the oracle.  TEST INFRASTRUCTURE ONLY:
// nothing in the product loads this library.
#include "../../nemo_cvpr2023_b200/csrc/rotmath.cuh"
#include "../../nemo_cvpr2023_b200/csrc/pathmath.cuh"

namespace nb { void set_error(const char*, ...) {} }
using namespace nb;

extern "C" {
void hc_rot6d(const float* x, float* R, int B) { for (int i = 0; i < B; ++i) rot6d_fwd(x + 6 * i, R + 9 * i); }
void hc_rot6d_bwd(const float* x, const float* dR, float* dx, int B) { for (int i = 0; i < B; ++i) rot6d_bwd(x + 6 * i, dR + 9 * i, dx + 6 * i); }
void hc_r2aa(const float* R, float* aa, int B, int nz) { for (int i = 0; i < B; ++i) rotmat2aa_fwd(R + 9 * i, aa + 3 * i, nz); }
void hc_r2aa_bwd(const float* R, const float* daa, float* dR, int B, int nz) { for (int i = 0; i < B; ++i) rotmat2aa_bwd(R + 9 * i, daa + 3 * i, dR + 9 * i, nz); }
void hc_rod(const float* aa, float* R, int B, int kind) { for (int i = 0; i < B; ++i) { if (kind == 0) rodrigues_quat_fwd(aa + 3 * i, R + 9 * i); else rodrigues_skew_fwd(aa + 3 * i, R + 9 * i); } }
void hc_rod_bwd(const float* aa, const float* dR, float* daa, int B, int kind) { for (int i = 0; i < B; ++i) { if (kind == 0) rodrigues_quat_bwd(aa + 3 * i, dR + 9 * i, daa + 3 * i); else rodrigues_skew_bwd(aa + 3 * i, dR + 9 * i, daa + 3 * i); } }
void hc_chain(const float* R, const float* J, int js, const int* parents, float* A, float* JP, int N) {
  float G[288];
  for (int n = 0; n < N; ++n) chain_fwd_frame(R + 216 * n, J + js * n, parents, G, A + 288 * n, JP + 72 * n);
}
void hc_chain_bwd(const float* R, const float* J, int js, const int* parents, const float* dA, const float* dJP, float* dR, float* dJ, int N) {
  float G[288], dG[288], At[288], JPt[72];
  for (int n = 0; n < N; ++n) {
    chain_fwd_frame(R + 216 * n, J + js * n, parents, G, At, JPt);
    chain_bwd_frame(R + 216 * n, J + js * n, parents, G, dA ? dA + 288 * n : nullptr, nullptr, dJP ? dJP + 72 * n : nullptr, dG, dR + 216 * n, dJ + 72 * n);
  }
}
void hc_project(const float* R, const float* t, float f, float cx, float cy, float eps, const float* p, float* uv, int P) {
  for (int i = 0; i < P; ++i) project_point(R, t, f, cx, cy, eps, p + 3 * i, uv + 2 * i, nullptr);
}
void hc_project_bwd(const float* R, const float* t, float f, float cx, float cy, float eps, const float* p, const float* duv, float* dp, float* dR, float* dt, int P) {
  for (int k = 0; k < 9; ++k) dR[k] = 0; for (int k = 0; k < 3; ++k) dt[k] = 0;
  for (int i = 0; i < P; ++i) {
    float uv[2], cam[3], dc[3];
    project_point(R, t, f, cx, cy, eps, p + 3 * i, uv, cam);
    project_point_bwd(cam, f, cx, cy, eps, duv + 2 * i, dc);
    for (int c = 0; c < 3; ++c) dp[3 * i + c] = R[c] * dc[0] + R[3 + c] * dc[1] + R[6 + c] * dc[2];
    for (int r = 0; r < 3; ++r) { dt[r] += dc[r]; for (int c = 0; c < 3; ++c) dR[3 * r + c] += dc[r] * p[3 * i + c]; }
  }
}
// out (n, ncomp), dr (n, 2) for weights w = 1
void hc_kploss(int mode, const float* pred, const float* gt, const float* conf, const float* size, int NP, int P, float rho, float* out, float* dr) {
  int nc = kl_ncomp(mode);
  float w[2] = {1.f, 1.f};
  for (int i = 0; i < NP; ++i)
    keypoint_loss_elem(mode, pred[2 * i] - gt[2 * i], pred[2 * i + 1] - gt[2 * i + 1], conf[i] > 0.5f ? 1.f : 0.f, size[i / P], rho, out + nc * i, w, dr + 2 * i);
}
void hc_rbf(int basis, const float* a, float* v, float* d, int n) { for (int i = 0; i < n; ++i) v[i] = rbf_basis(basis, a[i], d + i); }
void hc_adam(float* p, const float* g, float* m, float* v, int n, float lr, float b1, float b2, float eps, float wd, int dec, int step) {
  float bc1 = 1.f - powf(b1, (float)step), bc2s = sqrtf(1.f - powf(b2, (float)step));
  for (int i = 0; i < n; ++i) adam_elem(p + i, g[i], m + i, v + i, lr, b1, b2, eps, wd, dec, bc1, bc2s);
}
float hc_softplus(float x) { return softplusf_(x); }
}
