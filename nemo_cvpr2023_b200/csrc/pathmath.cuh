// Per-element math of the fit path (host+device so tests/hostcheck can pin it on the CPU):
// kinematic chain, projection, keypoint losses, phase warp, RBF bases, KL, Adam.
#pragma once
#include "common.cuh"

namespace nb {

// ------------------------------------------------------------------ kinematic chain
// batch_rigid_transform, human_body_prior/body_model/lbs.py:350-404.  One frame.
// R (24,9) row-major rotations, J (24,3) rest joints, parents[24] (parents[0] = -1).
// G (24,12): world 3x4 [R|t] per joint (scratch, kept for the backward).
NB_HD void chain_fwd_frame(const float* R, const float* J, const int* parents, float* G, float* A, float* JP) {
  for (int i = 0; i < NJ; ++i) {
    const float* Ri = R + 9 * i;
    float* Gi = G + 12 * i;
    int p = parents[i];
    if (i == 0 || p < 0) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        Gi[4 * r + 0] = Ri[3 * r]; Gi[4 * r + 1] = Ri[3 * r + 1]; Gi[4 * r + 2] = Ri[3 * r + 2];
        Gi[4 * r + 3] = J[3 * i + r];
      }
    } else {
      const float* Gp = G + 12 * p;
      float t[3] = {J[3 * i] - J[3 * p], J[3 * i + 1] - J[3 * p + 1], J[3 * i + 2] - J[3 * p + 2]};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        float g0 = Gp[4 * r], g1 = Gp[4 * r + 1], g2 = Gp[4 * r + 2];
        Gi[4 * r + 0] = g0 * Ri[0] + g1 * Ri[3] + g2 * Ri[6];
        Gi[4 * r + 1] = g0 * Ri[1] + g1 * Ri[4] + g2 * Ri[7];
        Gi[4 * r + 2] = g0 * Ri[2] + g1 * Ri[5] + g2 * Ri[8];
        Gi[4 * r + 3] = g0 * t[0] + g1 * t[1] + g2 * t[2] + Gp[4 * r + 3];
      }
    }
    float* Ai = A + 12 * i;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      Ai[4 * r] = Gi[4 * r]; Ai[4 * r + 1] = Gi[4 * r + 1]; Ai[4 * r + 2] = Gi[4 * r + 2];
      Ai[4 * r + 3] = Gi[4 * r + 3] - (Gi[4 * r] * J[3 * i] + Gi[4 * r + 1] * J[3 * i + 1] + Gi[4 * r + 2] * J[3 * i + 2]);
      JP[3 * i + r] = Gi[4 * r + 3];
    }
  }
}

// Backward.  G must hold the forward world transforms.  dG (24,12) scratch.
// dA (24,12) or NULL, dJP (24,3) or NULL -> dR (24,9), dJ (24,3) (both overwritten).
// dA (24,12) and dAt (12,24: the same gradient transposed, as the skinning backward reduces it) are summed; either may
// be NULL.
NB_HD void chain_bwd_frame(const float* R, const float* J, const int* parents, const float* G,
                           const float* dA, const float* dAt, const float* dJP, float* dG, float* dR, float* dJ) {
  for (int i = 0; i < NJ; ++i) {
    const float* Gi = G + 12 * i;
    float* dGi = dG + 12 * i;
    float at[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float a3 = (dA ? dA[12 * i + 4 * r + 3] : 0.f) + (dAt ? dAt[(4 * r + 3) * NJ + i] : 0.f);
      at[r] = a3;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        dGi[4 * r + c] = (dA ? dA[12 * i + 4 * r + c] : 0.f) + (dAt ? dAt[(4 * r + c) * NJ + i] : 0.f) - a3 * J[3 * i + c];
      dGi[4 * r + 3] = a3 + (dJP ? dJP[3 * i + r] : 0.f);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) dJ[3 * i + c] = -(Gi[c] * at[0] + Gi[4 + c] * at[1] + Gi[8 + c] * at[2]);
  }
  for (int i = NJ - 1; i >= 1; --i) {
    int p = parents[i];
    const float* Gp = G + 12 * p;
    const float* Ri = R + 9 * i;
    float* dGi = dG + 12 * i;
    float* dGp = dG + 12 * p;
    float t[3] = {J[3 * i] - J[3 * p], J[3 * i + 1] - J[3 * p + 1], J[3 * i + 2] - J[3 * p + 2]};
    // dR_i = Gp.R^T dG_i.R ; dt = Gp.R^T dG_i.t
    float dt[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int b = 0; b < 3; ++b)
        dR[9 * i + 3 * a + b] = Gp[a] * dGi[b] + Gp[4 + a] * dGi[4 + b] + Gp[8 + a] * dGi[8 + b];
      dt[a] = Gp[a] * dGi[3] + Gp[4 + a] * dGi[7] + Gp[8 + a] * dGi[11];
    }
    // dGp.R += dG_i.R R_i^T + dG_i.t (x) t ; dGp.t += dG_i.t
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        dGp[4 * r + c] += dGi[4 * r] * Ri[3 * c] + dGi[4 * r + 1] * Ri[3 * c + 1] + dGi[4 * r + 2] * Ri[3 * c + 2] +
                          dGi[4 * r + 3] * t[c];
      dGp[4 * r + 3] += dGi[4 * r + 3];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { dJ[3 * i + c] += dt[c]; dJ[3 * p + c] -= dt[c]; }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) dR[3 * r + c] = dG[4 * r + c];
    dJ[r] += dG[4 * r + 3];
  }
}

// ------------------------------------------------------------------ projection
// perspective_projection: hmr/geometry.py:78-106 (eps = 0), utils/geometry.py:63-93 (eps = 1e-8).
// K.(p/z) keeps the third component p_z/z, which multiplies the principal point.
NB_HD void project_point(const float* R, const float* t, float f, float cx, float cy, float eps,
                         const float* p, float* uv, float* cam /* 3: transformed point, may be NULL */) {
  float x = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + t[0];
  float y = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + t[1];
  float z = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + t[2];
  float zz = z + eps;
  float w = z / zz;
  uv[0] = f * (x / zz) + cx * w;
  uv[1] = f * (y / zz) + cy * w;
  if (cam) { cam[0] = x; cam[1] = y; cam[2] = z; }
}

// duv -> gradient wrt the transformed point (3)
NB_HD void project_point_bwd(const float* cam, float f, float cx, float cy, float eps, const float* duv, float* dcam) {
  float zz = cam[2] + eps;
  float iz = 1.f / zz;
  dcam[0] = duv[0] * f * iz;
  dcam[1] = duv[1] * f * iz;
  dcam[2] = -(duv[0] * f * cam[0] + duv[1] * f * cam[1]) * iz * iz + (cx * duv[0] + cy * duv[1]) * eps * iz * iz;
}

// ------------------------------------------------------------------ keypoint losses
// MultiViewModel.keypoint_loss, nemo/neural_motion_model.py:2806-2843; GMoF nemo/utils/misc_utils.py:91-105.
// r = pred - gt (2 components), m = (conf > 0.5).  Modes 0,1,3 give ONE value per joint (out[0]),
// modes 2,4,5 two.  dr = d(sum_k w[k]*out[k]) / d pred.
enum { KL_RMSE = 0, KL_RMSE_RESIZED = 1, KL_MSE = 2, KL_RMSE_ROBUST = 3, KL_MSE_ROBUST = 4, KL_MSE_ROBUST_RESIZED = 5 };
NB_HD int kl_ncomp(int mode) { return (mode == KL_RMSE || mode == KL_RMSE_RESIZED || mode == KL_RMSE_ROBUST) ? 1 : 2; }

NB_HD void keypoint_loss_elem(int mode, float rx, float ry, float m, float size, float rho, float* out,
                              const float* w, float* dr) {
  float rho2 = rho * rho;
  if (dr) { dr[0] = 0.f; dr[1] = 0.f; }
  if (mode == KL_RMSE || mode == KL_RMSE_RESIZED) {
    float s = (mode == KL_RMSE_RESIZED) ? 1.f / size : 1.f;
    float ax = rx * s, ay = ry * s;
    float q = sqrtf(1e-6f + ax * ax + ay * ay);
    out[0] = m * q;
    if (dr) { dr[0] = w[0] * m * ax / q * s; dr[1] = w[0] * m * ay / q * s; }
  } else if (mode == KL_MSE) {
    out[0] = m * rx * rx; out[1] = m * ry * ry;
    if (dr) { dr[0] = w[0] * m * 2.f * rx; dr[1] = w[1] * m * 2.f * ry; }
  } else if (mode == KL_RMSE_ROBUST) {
    float q = sqrtf(rx * rx + ry * ry);
    out[0] = m * rho2 * (q / (q + rho2));
    if (dr && q > 0.f) {
      float g = w[0] * m * rho2 * rho2 / ((q + rho2) * (q + rho2)) / q;
      dr[0] = g * rx; dr[1] = g * ry;
    }
  } else {
    float s = (mode == KL_MSE_ROBUST_RESIZED) ? 1000.f / size : 1.f;
    float ax = rx * s, ay = ry * s;
    float sx = ax * ax, sy = ay * ay;
    out[0] = m * rho2 * (sx / (sx + rho2));
    out[1] = m * rho2 * (sy / (sy + rho2));
    if (dr) {
      dr[0] = w[0] * m * rho2 * rho2 * 2.f * ax / ((sx + rho2) * (sx + rho2)) * s;
      dr[1] = w[1] * m * rho2 * rho2 * 2.f * ay / ((sy + rho2) * (sy + rho2)) * s;
    }
  }
}

// ------------------------------------------------------------------ phase warp + RBF
NB_HD float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

// RBF bases, nemo/rbf.py:62-118 (selected there with eval(name), :41): value and d/d alpha.
enum { RBF_GAUSSIAN = 0, RBF_LINEAR, RBF_QUADRATIC, RBF_INV_QUADRATIC, RBF_MULTIQUADRIC, RBF_INV_MULTIQUADRIC,
       RBF_SPLINE, RBF_POISSON_ONE, RBF_POISSON_TWO, RBF_MATERN32, RBF_MATERN52 };
NB_HD float rbf_basis(int b, float a, float* da) {
  const float s3 = 1.7320508075688772f, s5 = 2.23606797749979f;
  float v, d;
  switch (b) {
    case RBF_GAUSSIAN: v = expf(-a * a); d = -2.f * a * v; break;
    case RBF_LINEAR: v = a; d = 1.f; break;
    case RBF_QUADRATIC: v = a * a; d = 2.f * a; break;
    case RBF_INV_QUADRATIC: { float q = 1.f + a * a; v = 1.f / q; d = -2.f * a / (q * q); } break;
    case RBF_MULTIQUADRIC: { float q = sqrtf(1.f + a * a); v = q; d = a / q; } break;
    case RBF_INV_MULTIQUADRIC: { float q = 1.f + a * a; float r = sqrtf(q); v = 1.f / r; d = -a / (q * r); } break;
    case RBF_SPLINE: { float l = logf(a + 1.f); v = a * a * l; d = 2.f * a * l + a * a / (a + 1.f); } break;
    case RBF_POISSON_ONE: { float e = expf(-a); v = (a - 1.f) * e; d = (2.f - a) * e; } break;
    case RBF_POISSON_TWO: { float e = expf(-a); v = ((a - 2.f) / 2.f) * a * e; d = e * (2.f * a - 1.f - 0.5f * a * a); } break;
    case RBF_MATERN32: { float e = expf(-s3 * a); v = (1.f + s3 * a) * e; d = -3.f * a * e; } break;
    default: { float e = expf(-s5 * a); v = (1.f + s5 * a + (5.f / 3.f) * a * a) * e; d = -(5.f / 3.f) * a * (1.f + s5 * a) * e; } break;
  }
  if (da) *da = d;
  return v;
}

// VPoser posterior: sigma = softplus(h), KL(N(mu,sigma)||N(0,1)) = -log sigma + (sigma^2 + mu^2)/2 - 1/2
NB_HD float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // F.softplus, threshold 20

// torch.optim.Adam / AdamW single-tensor update (torch/optim/adam.py _single_tensor_adam, amsgrad off)
NB_HD void adam_elem(float* p, float g, float* m, float* v, float lr, float b1, float b2, float eps, float wd,
                     int decoupled, float bc1, float bc2_sqrt) {
  float pv = *p;
  if (decoupled) pv *= (1.f - lr * wd);
  else if (wd != 0.f) g += wd * pv;
  float mm = *m + (g - *m) * (1.f - b1);   // exp_avg.lerp_(grad, 1 - beta1)
  float vv = *v * b2 + (1.f - b2) * g * g;
  float denom = sqrtf(vv) / bc2_sqrt + eps;
  pv -= (lr / bc1) * (mm / denom);
  *p = pv; *m = mm; *v = vv;
}

}  // namespace nb
