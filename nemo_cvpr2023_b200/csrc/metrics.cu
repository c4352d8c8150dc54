// Evaluation metrics of the fitted motion, batched on the device.
// Reference arithmetic: nemo/utils/pose_utils.py:78-159 (compute_similarity_transform, reconstruction_error: MPJPE /
// MPVPE with or without Procrustes alignment) and the PCK@0.05 closure of MultiViewModel.eval_2d
// (nemo/neural_motion_model.py:524-532).  The reference loops over frames in numpy; here one CTA handles one item.
#include "common.cuh"

using namespace nb;

namespace {

constexpr int MT = 256;  // threads per item

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sums of NV_ values per thread; result valid in every thread
template <int NV_>
__device__ void block_sum(float (&v)[NV_], float* sh /* [MT/32][NV_] */, float* out /* [NV_] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV_; ++i) {
    const float s = warp_sum_f(v[i]);
    if (lane == 0) sh[warp * NV_ + i] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV_) {
    float s = 0.f;
    for (int w = 0; w < MT / 32; ++w) s += sh[w * NV_ + threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV_; ++i) v[i] = out[i];
  __syncthreads();
}

// err[b] = mean_p | S1[b,p] - S2[b,p] |      (reconstruction_error(pa=False) before its reduction over b)
__global__ void __launch_bounds__(MT) k_point_error(const float* __restrict__ S1, const float* __restrict__ S2, int P,
                                                    float* __restrict__ err) {
  __shared__ float sh[MT / 32], out[1];
  const float* a = S1 + (size_t)blockIdx.x * P * 3;
  const float* b = S2 + (size_t)blockIdx.x * P * 3;
  float v[1] = {0.f};
  for (int p = threadIdx.x; p < P; p += MT) {
    const float dx = a[3 * p] - b[3 * p], dy = a[3 * p + 1] - b[3 * p + 1], dz = a[3 * p + 2] - b[3 * p + 2];
    v[0] += sqrtf(dx * dx + dy * dy + dz * dz);
  }
  block_sum<1>(v, sh, out);
  if (threadIdx.x == 0) err[blockIdx.x] = v[0] / (float)P;
}

// Largest-eigenvalue eigenvector of a symmetric 4x4 matrix by cyclic Jacobi rotations (double: one thread per item).
__device__ void jacobi4(double (&Nm)[4][4], double (&q)[4]) {
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 24; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < 4; ++i)
      for (int j = i + 1; j < 4; ++j) off += Nm[i][j] * Nm[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < 3; ++p)
      for (int r = p + 1; r < 4; ++r) {
        if (fabs(Nm[p][r]) < 1e-300) continue;
        const double theta = (Nm[r][r] - Nm[p][p]) / (2.0 * Nm[p][r]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          const double a = Nm[k][p], b = Nm[k][r];
          Nm[k][p] = c * a - s * b;
          Nm[k][r] = s * a + c * b;
        }
        for (int k = 0; k < 4; ++k) {
          const double a = Nm[p][k], b = Nm[r][k];
          Nm[p][k] = c * a - s * b;
          Nm[r][k] = s * a + c * b;
        }
        for (int k = 0; k < 4; ++k) {
          const double a = V[k][p], b = V[k][r];
          V[k][p] = c * a - s * b;
          V[k][r] = s * a + c * b;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (Nm[i][i] > Nm[best][best]) best = i;
  for (int k = 0; k < 4; ++k) q[k] = V[k][best];
}

// Similarity Procrustes of S1 onto S2 (pose_utils.py:78-137): S1_hat = s R S1 + t with R the proper rotation that
// maximises trace(R K), K = sum_p x1 x2^T (the reference takes it from the SVD of K with the reflection fix; the
// same R is the rotation of the dominant eigenvector of Horn's 4x4 matrix of K).  srt (B,13) = [s, R row-major, t].
__global__ void __launch_bounds__(MT) k_procrustes(const float* __restrict__ S1, const float* __restrict__ S2, int P,
                                                   float* __restrict__ S1_hat, float* __restrict__ err,
                                                   float* __restrict__ srt) {
  __shared__ float sh[(MT / 32) * 10], out[10];
  __shared__ float tr[13];
  const float* a = S1 + (size_t)blockIdx.x * P * 3;
  const float* b = S2 + (size_t)blockIdx.x * P * 3;
  float m[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int p = threadIdx.x; p < P; p += MT) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      m[c] += a[3 * p + c];
      m[3 + c] += b[3 * p + c];
    }
  }
  block_sum<6>(m, sh, out);
#pragma unroll
  for (int c = 0; c < 6; ++c) m[c] /= (float)P;
  float k[10] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // K (9) and var1
  for (int p = threadIdx.x; p < P; p += MT) {
    const float x1[3] = {a[3 * p] - m[0], a[3 * p + 1] - m[1], a[3 * p + 2] - m[2]};
    const float x2[3] = {b[3 * p] - m[3], b[3 * p + 1] - m[4], b[3 * p + 2] - m[5]};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) k[3 * i + j] += x1[i] * x2[j];
      k[9] += x1[i] * x1[i];
    }
  }
  block_sum<10>(k, sh, out);
  if (threadIdx.x == 0) {
    const double Sxx = k[0], Sxy = k[1], Sxz = k[2], Syx = k[3], Syy = k[4], Syz = k[5], Szx = k[6], Szy = k[7], Szz = k[8];
    double Nm[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                       {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                       {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                       {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
    double q[4];
    jacobi4(Nm, q);
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z),     2 * (x * z + w * y),
                         2 * (x * y + w * z),     1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                         2 * (x * z - w * y),     2 * (y * z + w * x),     1 - 2 * (x * x + y * y)};
    double trRK = 0.0;  // trace(R K) = sum_ij R[i][j] K[j][i]
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) trRK += R[3 * i + j] * (double)k[3 * j + i];
    const double s = trRK / (double)k[9];
    tr[0] = (float)s;
    for (int i = 0; i < 9; ++i) tr[1 + i] = (float)R[i];
    for (int i = 0; i < 3; ++i)
      tr[10 + i] = (float)((double)m[3 + i] - s * (R[3 * i] * m[0] + R[3 * i + 1] * m[1] + R[3 * i + 2] * m[2]));
  }
  __syncthreads();
  if (srt && threadIdx.x < 13) srt[(size_t)blockIdx.x * 13 + threadIdx.x] = tr[threadIdx.x];
  float e[1] = {0.f};
  float* o = S1_hat ? S1_hat + (size_t)blockIdx.x * P * 3 : nullptr;
  for (int p = threadIdx.x; p < P; p += MT) {
    const float x0 = a[3 * p], x1 = a[3 * p + 1], x2 = a[3 * p + 2];
    float h[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) h[i] = tr[0] * (tr[1 + 3 * i] * x0 + tr[2 + 3 * i] * x1 + tr[3 + 3 * i] * x2) + tr[10 + i];
    if (o) {
      o[3 * p] = h[0];
      o[3 * p + 1] = h[1];
      o[3 * p + 2] = h[2];
    }
    const float dx = h[0] - b[3 * p], dy = h[1] - b[3 * p + 1], dz = h[2] - b[3 * p + 2];
    e[0] += sqrtf(dx * dx + dy * dy + dz * dz);
  }
  block_sum<1>(e, sh, out);
  if (err && threadIdx.x == 0) err[blockIdx.x] = e[0] / (float)P;
}

// PCK@0.05 of eval_2d (:524-532): out[0] += #(conf > 0.5 and sqrt(1e-6 + |pred - gt|^2) < 0.05 * gt_size[b]),
// out[1] += #(conf > 0.5)
__global__ void __launch_bounds__(MT) k_pck2d(const float* __restrict__ pred, const float* __restrict__ gt,
                                              const float* __restrict__ gt_size, int J, long long total,
                                              float* __restrict__ out2) {
  __shared__ float sh[(MT / 32) * 2], out[2];
  float v[2] = {0.f, 0.f};
  for (long long i = (long long)blockIdx.x * MT + threadIdx.x; i < total; i += (long long)gridDim.x * MT) {
    const long long bidx = i / J;
    const float dx = pred[2 * i] - gt[3 * i], dy = pred[2 * i + 1] - gt[3 * i + 1];
    const float rmse = sqrtf(1e-6f + (dx * dx + dy * dy));
    const float mask = gt[3 * i + 2] > 0.5f ? 1.f : 0.f;
    v[0] += mask * (rmse < 0.05f * gt_size[bidx] ? 1.f : 0.f);
    v[1] += mask;
  }
  block_sum<2>(v, sh, out);
  if (threadIdx.x < 2) atomicAdd(out2 + threadIdx.x, v[threadIdx.x]);
}

}  // namespace

extern "C" int nemo_b200_point_error(const float* S1, const float* S2, int P, int B, float* err, void* stream) {
  if (B == 0) return 0;
  NB_REQUIRE(S1 && S2 && err && P > 0 && B >= 0, "point_error: bad args");
  k_point_error<<<B, MT, 0, S(stream)>>>(S1, S2, P, err);
  NB_LAUNCH_CHECK("point_error");
}

extern "C" int nemo_b200_procrustes(const float* S1, const float* S2, int P, int B, float* S1_hat, float* err,
                                    float* srt, void* stream) {
  if (B == 0) return 0;
  NB_REQUIRE(S1 && S2 && P > 0 && B >= 0 && (S1_hat || err || srt), "procrustes: bad args");
  k_procrustes<<<B, MT, 0, S(stream)>>>(S1, S2, P, S1_hat, err, srt);
  NB_LAUNCH_CHECK("procrustes");
}

extern "C" int nemo_b200_pck2d(const float* pred, const float* gt, const float* gt_size, int J, int B, float* out2,
                               void* stream) {
  if (B == 0) return 0;
  NB_REQUIRE(pred && gt && gt_size && out2 && J > 0 && B >= 0, "pck2d: bad args");
  const long long total = (long long)B * J;
  k_pck2d<<<(int)min((total + MT - 1) / MT, (long long)1184), MT, 0, S(stream)>>>(pred, gt, gt_size, J, total, out2);
  NB_LAUNCH_CHECK("pck2d");
}
