// Device-side rotation math shared by the K3 kernels and the fused camera/loss kernel.
// Each function cites the reference code it reproduces (paths under /root/reference).
#pragma once
#include "common.cuh"

namespace nb {

// ---- 6D -> rotation matrix, hmr/geometry.py:47-61.  x = (x0..x5), a1 = (x0,x2,x4), a2 = (x1,x3,x5);
// R columns = (b1, b2, b3).  R is row-major R[3*i+j].
struct Rot6dCtx {
  float b1[3], b2[3], a2[3], n1, n2, d;
  bool c1, c2;  // norm above the 1e-12 clamp of F.normalize
};

NB_HD void rot6d_fwd(const float* x, float* R, Rot6dCtx* ctx = nullptr) {
  float a1[3] = {x[0], x[2], x[4]}, a2[3] = {x[1], x[3], x[5]};
  float n1 = sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]);
  bool c1 = n1 > 1e-12f;
  float m1 = c1 ? n1 : 1e-12f;
  float b1[3] = {a1[0] / m1, a1[1] / m1, a1[2] / m1};
  float d = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
  float u[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
  float n2 = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  bool c2 = n2 > 1e-12f;
  float m2 = c2 ? n2 : 1e-12f;
  float b2[3] = {u[0] / m2, u[1] / m2, u[2] / m2};
  float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    R[3 * i + 0] = b1[i];
    R[3 * i + 1] = b2[i];
    R[3 * i + 2] = b3[i];
  }
  if (ctx) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { ctx->b1[i] = b1[i]; ctx->b2[i] = b2[i]; ctx->a2[i] = a2[i]; }
    ctx->n1 = m1; ctx->n2 = m2; ctx->d = d; ctx->c1 = c1; ctx->c2 = c2;
  }
}

// backward of rot6d_fwd: dR (row-major 3x3) -> dx (6)
NB_HD void rot6d_bwd(const float* x, const float* dR, float* dx) {
  float R[9];
  Rot6dCtx c;
  rot6d_fwd(x, R, &c);
  float g1[3], g2[3], g3[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { g1[i] = dR[3 * i]; g2[i] = dR[3 * i + 1]; g3[i] = dR[3 * i + 2]; }
  const float *b1 = c.b1, *b2 = c.b2;
  // b3 = b1 x b2
  g1[0] += b2[1] * g3[2] - b2[2] * g3[1];
  g1[1] += b2[2] * g3[0] - b2[0] * g3[2];
  g1[2] += b2[0] * g3[1] - b2[1] * g3[0];
  g2[0] += g3[1] * b1[2] - g3[2] * b1[1];
  g2[1] += g3[2] * b1[0] - g3[0] * b1[2];
  g2[2] += g3[0] * b1[1] - g3[1] * b1[0];
  // b2 = u / max(|u|, eps)
  float du[3];
  if (c.c2) {
    float p = b2[0] * g2[0] + b2[1] * g2[1] + b2[2] * g2[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) du[i] = (g2[i] - b2[i] * p) / c.n2;
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) du[i] = g2[i] / c.n2;
  }
  // u = a2 - (b1.a2) b1
  float q = du[0] * b1[0] + du[1] * b1[1] + du[2] * b1[2];
  float da2[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    da2[i] = du[i] - q * b1[i];
    g1[i] += -q * c.a2[i] - c.d * du[i];
  }
  float da1[3];
  if (c.c1) {
    float p = b1[0] * g1[0] + b1[1] * g1[1] + b1[2] * g1[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) da1[i] = (g1[i] - b1[i] * p) / c.n1;
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) da1[i] = g1[i] / c.n1;
  }
  dx[0] = da1[0]; dx[2] = da1[1]; dx[4] = da1[2];
  dx[1] = da2[0]; dx[3] = da2[1]; dx[5] = da2[2];
}

// ---- rotation matrix -> axis angle, hmr/geometry.py:266-346 (rotation_matrix_to_quaternion, eps 1e-6)
// then :213-263 (quaternion_to_angle_axis).  Branch b in 0..3 selected by the diagonal.
NB_HD int r2q_select(const float* R, float* qs, float* t) {
  const float r00 = R[0], r01 = R[1], r02 = R[2], r10 = R[3], r11 = R[4], r12 = R[5], r20 = R[6], r21 = R[7],
              r22 = R[8];
  bool d2 = r22 < 1e-6f, d01 = r00 > r11, d0n1 = r00 < -r11;
  int b = d2 ? (d01 ? 0 : 1) : (d0n1 ? 2 : 3);
  if (b == 0) {
    *t = 1.f + r00 - r11 - r22;
    qs[0] = r21 - r12; qs[1] = *t; qs[2] = r10 + r01; qs[3] = r02 + r20;
  } else if (b == 1) {
    *t = 1.f - r00 + r11 - r22;
    qs[0] = r02 - r20; qs[1] = r10 + r01; qs[2] = *t; qs[3] = r21 + r12;
  } else if (b == 2) {
    *t = 1.f - r00 - r11 + r22;
    qs[0] = r10 - r01; qs[1] = r02 + r20; qs[2] = r21 + r12; qs[3] = *t;
  } else {
    *t = 1.f + r00 + r11 + r22;
    qs[0] = *t; qs[1] = r21 - r12; qs[2] = r02 - r20; qs[3] = r10 - r01;
  }
  return b;
}

NB_HD void rotmat2aa_fwd(const float* R, float* aa, bool nan_to_zero) {
  float qs[4], t;
  r2q_select(R, qs, &t);
  float is = 0.5f / sqrtf(t);
  float q0 = qs[0] * is, q1 = qs[1] * is, q2 = qs[2] * is, q3 = qs[3] * is;
  float s2 = q1 * q1 + q2 * q2 + q3 * q3;
  float s = sqrtf(s2);
  float tt = 2.f * (q0 < 0.f ? atan2f(-s, -q0) : atan2f(s, q0));
  float k = s2 > 0.f ? tt / s : 2.f;
  aa[0] = q1 * k; aa[1] = q2 * k; aa[2] = q3 * k;
  if (nan_to_zero) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (isnan(aa[i])) aa[i] = 0.f;
  }
}

// backward: daa (3) -> dR (9, row-major).  At the s2 == 0 singularity the reference's autograd
// yields NaN (0 * inf through the unselected torch.where branch); here the selected branch
// k = 2 is differentiated instead so no NaN is emitted.
NB_HD void rotmat2aa_bwd(const float* R, const float* daa_in, float* dR, bool nan_to_zero) {
  float qs[4], t;
  int b = r2q_select(R, qs, &t);
  float is = 0.5f / sqrtf(t);
  float q[4] = {qs[0] * is, qs[1] * is, qs[2] * is, qs[3] * is};
  float s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  float s = sqrtf(s2);
  float c = q[0];
  float tt = 2.f * (c < 0.f ? atan2f(-s, -c) : atan2f(s, c));
  float k = s2 > 0.f ? tt / s : 2.f;
  float daa[3] = {daa_in[0], daa_in[1], daa_in[2]};
  if (nan_to_zero) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (isnan(q[i + 1] * k)) daa[i] = 0.f;
  }
  float dq[4];
  dq[1] = daa[0] * k; dq[2] = daa[1] * k; dq[3] = daa[2] * k; dq[0] = 0.f;
  if (s2 > 0.f) {
    float dk = daa[0] * q[1] + daa[1] * q[2] + daa[2] * q[3];
    float dtt = dk / s;
    float ds = -dk * tt / s2;
    float r2 = s2 + c * c;
    ds += dtt * 2.f * c / r2;
    dq[0] = -dtt * 2.f * s / r2;
    float ds2 = ds / (2.f * s);
    dq[1] += 2.f * q[1] * ds2; dq[2] += 2.f * q[2] * ds2; dq[3] += 2.f * q[3] * ds2;
  }
  // q = 0.5 qs / sqrt(t)
  float dqs[4] = {dq[0] * is, dq[1] * is, dq[2] * is, dq[3] * is};
  float dt = -(dq[0] * q[0] + dq[1] * q[1] + dq[2] * q[2] + dq[3] * q[3]) / (2.f * t);
  float g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // r00 r01 r02 r10 r11 r12 r20 r21 r22
  if (b == 0) {
    dt += dqs[1];
    g[7] += dqs[0]; g[5] -= dqs[0]; g[3] += dqs[2]; g[1] += dqs[2]; g[2] += dqs[3]; g[6] += dqs[3];
    g[0] += dt; g[4] -= dt; g[8] -= dt;
  } else if (b == 1) {
    dt += dqs[2];
    g[2] += dqs[0]; g[6] -= dqs[0]; g[3] += dqs[1]; g[1] += dqs[1]; g[7] += dqs[3]; g[5] += dqs[3];
    g[0] -= dt; g[4] += dt; g[8] -= dt;
  } else if (b == 2) {
    dt += dqs[3];
    g[3] += dqs[0]; g[1] -= dqs[0]; g[2] += dqs[1]; g[6] += dqs[1]; g[7] += dqs[2]; g[5] += dqs[2];
    g[0] -= dt; g[4] -= dt; g[8] += dt;
  } else {
    dt += dqs[0];
    g[7] += dqs[1]; g[5] -= dqs[1]; g[2] += dqs[2]; g[6] -= dqs[2]; g[3] += dqs[3]; g[1] -= dqs[3];
    g[0] += dt; g[4] += dt; g[8] += dt;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) dR[i] = g[i];
}

// ---- axis angle -> rotation matrix, quaternion form, hmr/geometry.py:9-45
NB_HD void rodrigues_quat_fwd(const float* th, float* R, float* save = nullptr) {
  float e0 = th[0] + 1e-8f, e1 = th[1] + 1e-8f, e2 = th[2] + 1e-8f;
  float ang = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
  float ax[3] = {th[0] / ang, th[1] / ang, th[2] / ang};
  float sh, ch;
  sincosf(0.5f * ang, &sh, &ch);
  float qr[4] = {ch, sh * ax[0], sh * ax[1], sh * ax[2]};
  float qn = sqrtf(qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
  float w = qr[0] / qn, x = qr[1] / qn, y = qr[2] / qn, z = qr[3] / qn;
  float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz; R[2] = 2 * wy + 2 * xz;
  R[3] = 2 * wz + 2 * xy; R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
  R[6] = 2 * xz - 2 * wy; R[7] = 2 * wx + 2 * yz; R[8] = w2 - x2 - y2 + z2;
  if (save) {
    save[0] = w; save[1] = x; save[2] = y; save[3] = z; save[4] = qn; save[5] = ang; save[6] = sh; save[7] = ch;
  }
}

NB_HD void rodrigues_quat_bwd(const float* th, const float* dR, float* dth) {
  float R[9], sv[8];
  rodrigues_quat_fwd(th, R, sv);
  float w = sv[0], x = sv[1], y = sv[2], z = sv[3], qn = sv[4], ang = sv[5], sh = sv[6], ch = sv[7];
  float dn[4];
  dn[0] = 2 * w * (dR[0] + dR[4] + dR[8]) + 2 * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
  dn[1] = 2 * x * (dR[0] - dR[4] - dR[8]) + 2 * (y * dR[1] + z * dR[2] + y * dR[3] - w * dR[5] + z * dR[6] + w * dR[7]);
  dn[2] = 2 * y * (-dR[0] + dR[4] - dR[8]) + 2 * (x * dR[1] + w * dR[2] + x * dR[3] + z * dR[5] - w * dR[6] + z * dR[7]);
  dn[3] = 2 * z * (-dR[0] - dR[4] + dR[8]) + 2 * (-w * dR[1] + x * dR[2] + w * dR[3] + y * dR[5] + x * dR[6] + y * dR[7]);
  float p = w * dn[0] + x * dn[1] + y * dn[2] + z * dn[3];
  float dq[4] = {(dn[0] - w * p) / qn, (dn[1] - x * p) / qn, (dn[2] - y * p) / qn, (dn[3] - z * p) / qn};
  float ax[3] = {th[0] / ang, th[1] / ang, th[2] / ang};
  float dh = -sh * dq[0] + ch * (dq[1] * ax[0] + dq[2] * ax[1] + dq[3] * ax[2]);
  float dax[3] = {sh * dq[1], sh * dq[2], sh * dq[3]};
  float dang = 0.5f * dh - (dax[0] * th[0] + dax[1] * th[1] + dax[2] * th[2]) / (ang * ang);
#pragma unroll
  for (int i = 0; i < 3; ++i) dth[i] = dax[i] / ang + dang * (th[i] + 1e-8f) / ang;
}

// ---- axis angle -> rotation matrix, skew form, human_body_prior/body_model/lbs.py:303-334
NB_HD void rodrigues_skew_fwd(const float* r, float* R) {
  float e0 = r[0] + 1e-8f, e1 = r[1] + 1e-8f, e2 = r[2] + 1e-8f;
  float ang = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
  float dx = r[0] / ang, dy = r[1] / ang, dz = r[2] / ang;
  float s, c;
  sincosf(ang, &s, &c);
  float K[9] = {0, -dz, dy, dz, 0, -dx, -dy, dx, 0};
  float K2[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) K2[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.f : 0.f) + s * K[i] + (1.f - c) * K2[i];
}

NB_HD void rodrigues_skew_bwd(const float* r, const float* dR, float* dr) {
  float e0 = r[0] + 1e-8f, e1 = r[1] + 1e-8f, e2 = r[2] + 1e-8f;
  float ang = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
  float dx = r[0] / ang, dy = r[1] / ang, dz = r[2] / ang;
  float s, c;
  sincosf(ang, &s, &c);
  float K[9] = {0, -dz, dy, dz, 0, -dx, -dy, dx, 0};
  float K2[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) K2[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
  float gs = 0.f, gc = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) { gs += dR[i] * K[i]; gc -= dR[i] * K2[i]; }
  // dK = s dR + (1-c) (dR K^T + K^T dR)
  float dK[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a = 0.f;
#pragma unroll
      for (int m = 0; m < 3; ++m) a += dR[3 * i + m] * K[3 * j + m] + K[3 * m + i] * dR[3 * m + j];
      dK[3 * i + j] = s * dR[3 * i + j] + (1.f - c) * a;
    }
  float dd[3] = {dK[7] - dK[5], dK[2] - dK[6], dK[3] - dK[1]};
  float da = gs * c - gc * s;
  float dang = da - (dd[0] * r[0] + dd[1] * r[1] + dd[2] * r[2]) / (ang * ang);
  dr[0] = dd[0] / ang + dang * e0 / ang;
  dr[1] = dd[1] / ang + dang * e1 / ang;
  dr[2] = dd[2] / ang + dang * e2 / ang;
}

}  // namespace nb
