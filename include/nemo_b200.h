/* nemo_b200.h -- C ABI of libnemo_b200.so: the NeMo per-iteration fitting hot path on B200 (sm_100a).
 *
 * The reference (wangkua1/nemo-cvpr2023) is pure Python/PyTorch and defines no FFI for this path
 * (SURVEY.md section 8(b)); its boundary is Python call sites.  This header is the C-ABI that the
 * Python mirror of those call sites (nemo_cvpr2023_b200/) binds through ctypes; every entry point
 * names the reference function (file:line under /root/reference) whose arithmetic it replaces.
 *
 * Conventions (all entry points):
 *   - returns 0 on success, a negative NEMO_B200_E* code otherwise; text via nemo_b200_last_error()
 *   - every tensor argument is a DEVICE pointer to contiguous row-major fp32 unless stated
 *     (indices are int64, small tables int32); explicit sizes; the CALLER owns every buffer
 *   - no allocation, no device synchronisation, no host reads of device data inside
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); safe under CUDA-graph capture
 *   - "planar" vertex layout = [3][VP] per frame with VP = NEMO_B200_VP padded vertices (pad = 0)
 */
#ifndef NEMO_B200_H_
#define NEMO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEMO_B200_NV 6890   /* SMPL vertices */
#define NEMO_B200_VP 6912   /* padded to 27 x 256 */
#define NEMO_B200_NJ 24     /* kinematic joints */

#define NEMO_B200_EINVAL  (-1)  /* bad argument */
#define NEMO_B200_ECUDA   (-2)  /* CUDA runtime error at launch */
#define NEMO_B200_EARCH   (-3)  /* device is not sm_100 */

const char* nemo_b200_last_error(void);
int nemo_b200_version(void);
/* compute capability (major*10+minor) of the current device, or a negative error code */
int nemo_b200_sm_arch(void);

/* ---------------------------------------------------------------- GEMM (K2 / K5)
 * C[M,N] = epilogue( op(A)[M,K] . op(B)[K,N] ),  replaces the cuBLAS calls behind
 *   nn.Linear in FCNN/MotionNet (nemo/neural_motion_model.py:58-71,130-148), VPoser
 *   (human_body_prior/models/vposer_model.py:63-88), the pose-blend matmul (lbs.py:229-233),
 *   blend_shapes (lbs.py:277-300) and their autograd backward.
 * transA: A stored (K,M) row-major instead of (M,K).  transB: B stored (N,K) instead of (K,N).
 * epilogue: v = acc + bias[n] (bias may be NULL); act 0 none / 1 relu / 2 leaky_relu(0.01);
 *   if gate != NULL: v *= (gate[m*ldg+n] > 0 ? 1 : gate_slope)   (activation backward)
 *   if accumulate: C += v else C = v. */
int nemo_b200_gemm(const float* A, const float* B, const float* bias, float* C,
                   int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
                   int act, const float* gate, int ldg, float gate_slope, int accumulate,
                   void* stream);
/* nemo_b200_gemm dispatches: 16-byte aligned operands with lda, ldb multiples of 4 and M*N*K >= 2e6 run on a
 * tensor-core kernel, everything else on the CUDA-core kernel.  Row padding: when C is 16-byte addressable (aligned,
 * ldc % 4 == 0) and N % 4 != 0 with ldc - N < 4, the tensor-core kernels write the tile through 16-byte TMA stores and the
 * up to three PADDING floats [N, ldc) of each row receive zeros. */
/* same contract, always on the tcgen05 3xTF32 tensor-core kernel (TMA-fed; A and B must be 16-byte aligned
 * with lda, ldb multiples of 4).  splits > 0 forces the split-K factor (act must be 0), 0 = heuristic. */
int nemo_b200_gemm_tc(const float* A, const float* B, const float* bias, float* C,
                      int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
                      int act, const float* gate, int ldg, float gate_slope, int accumulate,
                      int splits, void* stream);
/* same contract on the second-generation tensor-core kernel (128 x 256 x 16 tiles, N = 256 MMAs, TMA epilogue only:
 * C and gate must additionally be 16-byte aligned with ldc, ldg multiples of 4, and N % 4 == 0 or ldc - N < 4).  nemo_b200_gemm picks between the two
 * by the estimated tensor-pipe time of the shape. */
int nemo_b200_gemm_tc2(const float* A, const float* B, const float* bias, float* C,
                       int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
                       int act, const float* gate, int ldg, float gate_slope, int accumulate,
                       int splits, void* stream);
/* C[M,N] = out_scale * ( (a_scale*A)[M,K] . (Bh + Bl)^T ) + bias   -- the pose-blend matmul of lbs (lbs.py:229-233) on fp16
 * operand PAIRS (tcgen05.mma.kind::f16, K = 16: half the MMAs of 3xTF32 for the same 22 mantissa bits).  A (M, lda) fp32,
 * scaled by the power of two a_scale and split on the fly; Bh, Bl (N, ldb) fp16, K-major: the hi / lo parts of
 * b_scale * B^T, split ONCE by the caller (B = posedirs is a constant of the body model); out_scale = 1 / (a_scale *
 * b_scale).  The caller guarantees |a_scale * A| < 65504 (the pose feature R - I lies in [-2, 2]: a_scale = 2^13).
 * All pointers 16-byte aligned, lda % 4 == 0, ldb % 8 == 0, ldc % 4 == 0, N % 4 == 0 or ldc - N < 4. */
int nemo_b200_gemm_h16(const float* A, const void* Bh, const void* Bl, const float* bias, float* C,
                       int M, int N, int K, int lda, int ldb, int ldc, float a_scale, float out_scale,
                       void* stream);
/* same contract, always on the CUDA cores (fp32 FMA); the cross-check of the tensor-core path */
int nemo_b200_gemm_simt(const float* A, const float* B, const float* bias, float* C,
                        int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
                        int act, const float* gate, int ldg, float gate_slope, int accumulate,
                        void* stream);
/* out[n] (+)= sum_m X[m*ld+n]   (bias gradients) */
int nemo_b200_colsum(const float* X, float* out, int M, int N, int ld, int accumulate, void* stream);

/* ---------------------------------------------------------------- rotations (K3)
 * Strided addressing (all six entry points): rotation i of B sits in row n = i / per_row, slot
 * j = i % per_row; a tensor with row stride ld (floats) and w floats per rotation is addressed at
 * base + n*ld + w*j.  Contiguous (B,w): per_row = 1, ld = w.  This lets the kernels run in place on
 * slices, e.g. rot6d = Y[:, :144] of the (N,147) MLP output or poses = aa[:, 3:72].
 * rot6d: hmr/geometry.py:47-61 and vposer_model.py:32-45. */
int nemo_b200_rot6d_fwd(const float* x6, int ld6, float* R, int ldr, int per_row, int B, void* stream);
int nemo_b200_rot6d_bwd(const float* x6, int ld6, const float* dR, int ldr, float* dx6, int ldd,
                        int per_row, int accumulate, int B, void* stream);
/* rotation matrix -> axis angle: hmr/geometry.py:181-346 (nan_to_zero=1) and
 * human_body_prior/tools/tgm_conversion.py:219-376 (nan_to_zero=0). */
int nemo_b200_rotmat2aa_fwd(const float* R, int ldr, float* aa, int lda, int per_row, int B,
                            int nan_to_zero, void* stream);
int nemo_b200_rotmat2aa_bwd(const float* R, int ldr, const float* daa, int lda, float* dR, int ldd,
                            int per_row, int accumulate, int B, int nan_to_zero, void* stream);
/* axis angle -> rotation matrix. kind 0: quaternion form, hmr/geometry.py:9-45.
 * kind 1: skew-matrix form, human_body_prior/body_model/lbs.py:303-334. */
int nemo_b200_rodrigues_fwd(const float* aa, int lda, float* R, int ldr, int per_row, int B, int kind,
                            void* stream);
int nemo_b200_rodrigues_bwd(const float* aa, int lda, const float* dR, int ldr, float* daa, int ldd,
                            int per_row, int accumulate, int B, int kind, void* stream);
/* pose feature (R[1:] - I) of lbs.py:221-228 written into F[n*ldf + 0..206]; R is (N,24,9). */
int nemo_b200_pose_feature(const float* R, float* F, int ldf, int N, void* stream);

/* ---------------------------------------------------------------- kinematic chain (K6)
 * batch_rigid_transform + transform_mat, lbs.py:337-404.  One warp per frame (lane = joint), the tree walked by depth;
 * any tree (parent < 0 = root).
 * R (N,24,9); J (24,3) shared (j_stride 0) or (N,24,3) (j_stride 72); parents int32[24].
 * out: A (N,24,12) relative 3x4 transforms, JP (N,24,3) posed joints. */
int nemo_b200_chain_fwd(const float* R, const float* J, int j_stride, const int* parents,
                        float* A, float* JP, int N, void* stream);
/* chain_fwd fused with what precedes it for one mesh (one launch instead of four): optionally the Rodrigues map of
 * the 23 body joints (aa != NULL: rows of `lda` floats, 69 used; kind as in nemo_b200_rodrigues_fwd; the root rotation is
 * copied from R0 (N,24,9), joint 0; all 24 rotations are written to R_out (N,24,9) if not NULL) -- else R (N,24,9) is
 * the input -- and optionally the pose feature (R[1:] - I) into F[n*ldf + 0..206] (lbs.py:221-228).  Exactly one of R
 * and aa is given.  One warp per frame, lane = joint. */
int nemo_b200_pose_chain_fwd(const float* R, const float* aa, int lda, int kind, const float* R0,
                             const float* J, int j_stride, const int* parents,
                             float* R_out, float* F, int ldf, float* A, float* JP, int N, void* stream);
/* dA (N,24,12) and dAt (N,12,24) (the same gradient transposed per frame) are summed;
 * either may be NULL.  dR (N,24,9) overwritten; dJ (N,24,3) overwritten when j_stride != 0, else (24,3)
 * ACCUMULATED with atomics (caller zeroes); dJ may be NULL (not computed). */
int nemo_b200_chain_bwd(const float* R, const float* J, int j_stride, const int* parents,
                        const float* dA, const float* dAt, const float* dJP, float* dR, float* dJ, int N,
                        void* stream);
/* chain_bwd + the backward of the pose feature: dF (N, ldf >= 207), the gradient of (R[1:] - I) -- what the dgrad of the
 * pose-blend GEMM produces, written to its own 16-byte aligned rows so that GEMM runs with the TMA epilogue -- is added
 * into dR[:, 1:].  dF may be NULL. */
int nemo_b200_chain_bwd2(const float* R, const float* J, int j_stride, const int* parents,
                         const float* dA, const float* dAt, const float* dJP, const float* dF, int ldf,
                         float* dR, float* dJ, int N, void* stream);

/* ---------------------------------------------------------------- skinning (K7)
 * lbs.py:236-252 (v_posed, T = W.A, v = T.[v_posed;1]), fused; no per-vertex intermediate is written.
 * v_posed[n] = base + PO[n].
 *   A    (N,24,12)           PO   (N,3,VP) planar      base (3,VP) planar
 *   Wt   (24,VP) transposed skinning weights, zero padded
 * out: V (N,6890,3) interleaved (the reference layout). */
int nemo_b200_skin_fwd(const float* A, const float* PO, const float* base, const float* Wt,
                       float* V, int N, void* stream);
/* backward.  Upstream gradient of the vertices = dV (N,6890,3) [may be NULL]
 *   + sign_scale * sign bits [signs may be NULL]; at least one of the two.
 * out: dPO (N,3,VP) overwritten, dA (N,24,12) overwritten (dA[n][j][r*4+c], the layout of A;
 * nemo_b200_chain_bwd consumes it as is).  d base = column sum of dPO. */
int nemo_b200_skin_bwd(const float* A, const float* PO, const float* base, const float* Wt,
                       const float* dV, const uint32_t* signs, float sign_scale,
                       float* dPO, float* dA, int N, void* stream);
/* fused vertex-to-vertex L1 of MultiViewModel.vposer_loss (nemo/neural_motion_model.py:2792-2797):
 * l1_sum[0] += sum_{n,v,c} | v_b - v_a |  (mesh a = "orig", carries the gradient; mesh b = detached
 * reconstruction).  signs (N,3,2,VP/32) uint32 receives the (v_a > v_b, v_a < v_b) bit masks that
 * nemo_b200_skin_bwd consumes, so the backward never recomputes mesh b. */
int nemo_b200_skin_l1_fwd(const float* A_a, const float* PO_a, const float* A_b, const float* PO_b,
                          const float* base, const float* Wt, uint32_t* signs, float* l1_sum, int N,
                          void* stream);

/* ---------------------------------------------------------------- evaluation metrics
 * nemo/utils/pose_utils.py:78-159 (compute_similarity_transform[_batch], reconstruction_error), batched: one CTA per
 * item instead of a numpy loop over frames.  S1, S2 (B,P,3).
 * point_error:  err[b] = mean_p |S1[b,p] - S2[b,p]|                       (reconstruction_error(pa=False))
 * procrustes:   S1_hat = s R S1 + t, the similarity transform that takes S1 closest to S2 (proper rotation: the
 *               reference's SVD solution with its reflection fix); err[b] = mean_p |S1_hat - S2|;
 *               srt (B,13) = [s, R row-major, t].  S1_hat, err, srt may each be NULL (not all three).
 * pck2d:        PCK@0.05 of MultiViewModel.eval_2d (nemo/neural_motion_model.py:524-532): pred (B,J,2),
 *               gt (B,J,3) = x, y, conf, gt_size (B); out2[0] += hits, out2[1] += valid (caller zeroes). */
int nemo_b200_point_error(const float* S1, const float* S2, int P, int B, float* err, void* stream);
int nemo_b200_procrustes(const float* S1, const float* S2, int P, int B, float* S1_hat, float* err, float* srt,
                         void* stream);
int nemo_b200_pck2d(const float* pred, const float* gt, const float* gt_size, int J, int B, float* out2,
                    void* stream);

/* ---------------------------------------------------------------- joint read-outs (K7j)
 * smplx VertexJointSelector (picked vertices) and hmr/smpl.py:33-36 (J_regressor_extra . vertices),
 * evaluated WITHOUT the mesh: a read-out is a fixed linear functional of the vertices and skinning is
 * linear in [v_posed;1] per joint, so   joint_q[n] = sum_j A[n,j]_3x3 . z[n,q,j] + A[n,j]_t * s[q,j]
 * with z[n,q,j] = sum_v w_q[v] W[v,j] v_posed[n,v] = a row of  Z = [pose feature | betas] . MZ + Z0
 * (one small GEMM; MZ is posedirs / shapedirs contracted with w_q (x) W once per asset set).
 *   Z  (N, ldz): [np picked vertices x 3 (their own v_posed) | nx regressed joints x 24 x 3]
 *   Wp (np,24) skinning weights of the picked vertices;  Sj (nx,24) = sum_v w_q[v] W[v,j]
 * out: JX (N, np+nx, 3): picked vertices then regressed joints (no translation). */
int nemo_b200_readout_fwd(const float* A, const float* Z, int ldz, const float* Wp, int np,
                          const float* Sj, int nx, float* JX, int N, void* stream);
/* dJX (N,np+nx,3) -> dA (N,24,12) (overwritten, or added to when accumulate != 0) and dZ (N, ldz)
 * overwritten (pad columns zeroed). */
int nemo_b200_readout_bwd(const float* A, const float* Z, int ldz, const float* Wp, int np,
                          const float* Sj, int nx, const float* dJX, float* dA, int accumulate,
                          float* dZ, int N, void* stream);

/* joints49 = cat(JP, JX)[joint_map] (+ transl), hmr/smpl.py:33-37 + smplx SMPL.forward.
 * JX rows: np picked vertices then nx regressed joints; regressed joints receive rowsum[r]*transl
 * (they are regressed from translated vertices), all others transl.  transl may be NULL. */
int nemo_b200_joints_gather_fwd(const float* JP, const float* JX, int nraw_x, int np,
                                const int* joint_map, int nout, const float* transl,
                                const float* rowsum, float* out, int N, void* stream);
/* dJP (N,24,3), dJX (N,nraw_x,3) overwritten; dtransl (N,3) overwritten if not NULL. */
int nemo_b200_joints_gather_bwd(const float* dout, int nraw_x, int np, const int* joint_map, int nout,
                                const float* rowsum, float* dJP, float* dJX, float* dtransl, int N,
                                void* stream);

/* ---------------------------------------------------------------- projection + 2D loss (K7b)
 * perspective_projection, hmr/geometry.py:78-106 (eps 0) / utils/geometry.py:63-93 (eps 1e-8).
 * points (B,P,3), rot (B,9), trans (B,3), focal (B), centre (B,2) -> out (B,P,2). */
int nemo_b200_project_fwd(const float* points, const float* rot, const float* trans, const float* focal,
                          const float* centre, float eps, float* out, int B, int P, void* stream);
int nemo_b200_project_bwd(const float* points, const float* rot, const float* trans, const float* focal,
                          float eps, const float* dout, float* dpoints, float* drot, float* dtrans,
                          int B, int P, void* stream);

/* learned_camera_projection (nemo/neural_motion_model.py:3073-3124) + keypoint_loss (:2806-2843,
 * GMoF nemo/utils/misc_utils.py:91-105) + the per-view reduction of step() (:3551-3558), fused.
 *   joints (N,P,3) already translated; cameras (V,9) = [t(3), rot6d(6)]; gt (V,F,P,3) = x,y,conf;
 *   bbox (V,F); loss_type 0 rmse 1 rmse_resized 2 mse 3 rmse_robust 4 mse_robust 5 mse_robust_resized
 *   reduction 0: step() -- mean over present views of mean(loss * conf) (:3551-3558);
 *   reduction 1: camera_fitting_loss -- plain mean over all elements (:2845-2867)
 * fwd: points2d (N,P,2), loss_all (N,P,2) [rmse* modes: (N,P,1) stored in the first N*P floats],
 *      view_acc (V,2) += {sum(loss*conf), sample count}   (caller zeroes)  */
int nemo_b200_kp_loss_fwd(const float* joints, const int64_t* view_idx, const int64_t* frame_idx,
                          const float* cameras, const float* gt, const float* bbox,
                          float focal, float cx, float cy, float rho, int loss_type, int reduction,
                          float* points2d, float* loss_all, float* view_acc,
                          int N, int P, int V, int F, void* stream);
/* kp = mean over present views of view_sum/(count*P*comp);  dloss = upstream d(kp) (device scalar
 * pointer, or NULL for 1.0) times `weight`.  Outputs: djoints (N,P,3) overwritten,
 * dcam_acc (V,12) += {dt(3), dRcam(9)} (caller zeroes), loss_out[0] = kp (may be NULL). */
int nemo_b200_kp_loss_bwd(const float* joints, const int64_t* view_idx, const int64_t* frame_idx,
                          const float* cameras, const float* gt, const float* bbox,
                          float focal, float cx, float cy, float rho, int loss_type, int reduction,
                          const float* view_acc, float weight,
                          float* djoints, float* dcam_acc, float* loss_out,
                          int N, int P, int V, int F, void* stream);
/* dcameras (V,9) (+)= rot6d backward of the accumulated camera gradients */
int nemo_b200_cam_finish(const float* cameras, const float* dcam_acc, float* dcameras, int accumulate,
                         int V, void* stream);

/* masked GMoF pose loss: mean(1[m>0.5] * rho^2 r^2/(r^2+rho^2)), r = pose - target
 * (3D loss nemo/neural_motion_model.py:3870-3882, warmup :3489-3491).  target/mask gathered by
 * (view_idx, frame_idx) from (V,F,D) / (V,F,1).  loss_acc[0] += weight-free mean; dpose (+)= weight*grad */
int nemo_b200_pose3d_loss(const float* pose, int ldp, const int64_t* view_idx, const int64_t* frame_idx,
                          const float* target, const float* mask, float rho, float weight,
                          float* loss_acc, float* dpose, int ldd, int accumulate,
                          int N, int D, int F, void* stream);

/* ---------------------------------------------------------------- motion-field input (K1)
 * frame_idx_to_raw_phase (:2978-2984) -> MonotonicNetwork.forward (monotonic_network.py:23-39) of the
 * sample's own view -> RBF.forward (nemo/rbf.py:47-56) -> cat(instance code) (:3741-3744).
 *   raw_phase (N) fp32; shifts, scales (V,K); centres, log_sigmas (D) (D = 0: raw warped phase is the
 *   feature); codes (V,C) (C may be 0); noise (N,C) or NULL.
 * out: X (N, ldx) rows [rbf(D or 1) | code(C)], warped (N) the warped phase, yn (V,3) = {y(0), y(1), den}. */
int nemo_b200_embed_fwd(const float* raw_phase, const int64_t* view_idx,
                        const float* shifts, const float* scales, int K,
                        const float* centres, const float* log_sigmas, int D, int basis,
                        const float* codes, int C, const float* noise,
                        float* X, int ldx, float* warped, float* yn, int N, int V, void* stream);
/* dX (N, ldx) -> dlog_sigmas (D), dshifts, dscales (V,K), dcodes (V,C): all ACCUMULATED (atomics). */
int nemo_b200_embed_bwd(const float* raw_phase, const int64_t* view_idx,
                        const float* shifts, const float* scales, int K,
                        const float* centres, const float* log_sigmas, int D, int basis,
                        int C, const float* warped, const float* yn, const float* dX, int ldx,
                        float* dlog_sigmas, float* dshifts, float* dscales, float* dcodes,
                        float* scratch /* (V,2) zeroed by the callee */, int N, int V, void* stream);
/* standalone layers with the reference module semantics */
int nemo_b200_rbf_fwd(const float* x, const float* centres, const float* log_sigmas, int D, int basis,
                      float* out, int N, void* stream);
int nemo_b200_rbf_bwd(const float* x, const float* centres, const float* log_sigmas, int D, int basis,
                      const float* dout, float* dx, float* dlog_sigmas /* accumulated */, int N,
                      void* stream);
int nemo_b200_monotonic_fwd(const float* x, const float* shifts, const float* scales, int K,
                            float* out, int N, void* stream);
int nemo_b200_monotonic_bwd(const float* x, const float* shifts, const float* scales, int K,
                            const float* dout, float* dx, float* dshifts, float* dscales /* accumulated */,
                            float* scratch /* 2 floats, zeroed by the callee */, int N, void* stream);

/* ---------------------------------------------------------------- priors
 * MaxMixturePrior.merged_log_likelihood, hmr/smplify/prior.py:181-196.  pose rows stride ldp.
 * ll (N) overwritten; if dpose != NULL: dpose[n] (+)= gscale * d ll[n] / d pose[n]. */
int nemo_b200_gmm_prior(const float* pose, int ldp, const float* means, const float* precisions,
                        const float* nll_weights, int M, int D, float* ll,
                        float gscale, float* dpose, int ldd, int accumulate, int N, void* stream);
/* The precisions are frozen buffers: nemo_b200_gmm_prepare symmetrises them once into prec_sym (M, 69, 72) (rows padded
 * to 16 bytes, 0.5 (P + P^T)) and nemo_b200_gmm_prior_sym is nemo_b200_gmm_prior reading that table. */
int nemo_b200_gmm_prepare(const float* precisions, int M, int D, float* prec_sym, void* stream);
int nemo_b200_gmm_prior_sym(const float* pose, int ldp, const float* means, const float* prec_sym,
                            const float* nll_weights, int M, int D, float* ll,
                            float gscale, float* dpose, int ldd, int accumulate, int N, void* stream);
/* VPoser posterior head (vposer_model.py:48-56) + KL(N(mu,sigma)||N(0,1)) of vposer_loss (:2798-2803).
 * head (N,2L) = [mu | pre-softplus]; kl_acc[0] += sum_n sum_d KL; sigma (N,L) out (may be NULL);
 * dhead (N,2L) = gscale * d(sum KL)/d head (may be NULL). */
int nemo_b200_vposer_kl(const float* head, int L, float* sigma, float* kl_acc, float gscale,
                        float* dhead, int N, void* stream);

/* ---------------------------------------------------------------- optimiser
 * torch.optim.Adam / AdamW single-tensor semantics (the reference's optimisers,
 * nemo/neural_motion_model.py:3404-3431) over one flat parameter buffer.  `step` is the 1-based count. */
int nemo_b200_adam(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int decoupled, int step,
                   const int* step_dev /* device-resident step count, overrides `step` when not NULL */,
                   void* stream);
int nemo_b200_increment(int* counter, void* stream);   /* *counter += 1 on the device */

/* Every optimiser of a fit model in ONE launch (the reference steps 3-4 torch.optim.Adam objects per iteration,
 * nemo/neural_motion_model.py:3586-3588, each a Python loop over parameter tensors).  A segment = a 16-byte aligned
 * range of a flat parameter buffer with its own hyper-parameters and a DEVICE count of completed steps (`step`;
 * segments may share one counter, exactly one of them carries OWNS_STEP).  The kernel uses *step + 1 as the step
 * number and advances the counters itself (no increment kernel; safe under CUDA-graph replay).  SKIP leaves the
 * segment and its counter untouched: torch.optim.Adam skips parameters whose .grad is None (the warm-up phase never
 * reaches linear_out.*, :3455-3509).  `segs` is a HOST array, copied into the launch parameters; `ticket` is a device
 * int that must be 0 on entry (the kernel resets it). */
#define NEMO_B200_ADAM_MAX_SEGS 16
#define NEMO_B200_ADAM_SKIP 1
#define NEMO_B200_ADAM_OWNS_STEP 2
typedef struct nemo_b200_adam_seg {
  float* p; const float* g; float* m; float* v;
  unsigned long long n;
  float lr, beta1, beta2, eps, weight_decay;
  int decoupled;   /* 0: torch.optim.Adam (L2 in the gradient), 1: AdamW */
  int* step;
  int flags;
} nemo_b200_adam_seg;
int nemo_b200_adam_multi(const nemo_b200_adam_seg* segs, int nseg, int* ticket, void* stream);

/* ---------------------------------------------------------------- data-parallel gradient sum over peer memory
 * The reference back-propagates ONE loss over all video instances (nemo/neural_motion_model.py:3511-3598); with the
 * instances sharded over the GPUs of a node the shared MotionNet + RBF gradient is the sum of the ranks' (SURVEY.md
 * 8(e)).  bufs: HOST array of `world` device pointers, entry r = rank r's symmetric buffer as mapped into this process
 * (n data floats followed, at float offset sig_off >= n, by a 64-word signal area zeroed once before the first call);
 * src / dst: the local gradient (may alias).  Three kernels on `stream` (capturable): stage, reduce slice `rank` of
 * every buffer in rank order and write it to every buffer, copy out.  Every rank of the group must make the same calls. */
int nemo_b200_p2p_allreduce(const void* const* bufs, int rank, int world, const float* src, float* dst, size_t n,
                            size_t sig_off, void* stream);

/* small utilities used by the host mirror */
/* out = dy * act'(y) for y = act(.), act 1 relu / 2 leaky_relu(0.01) */
int nemo_b200_act_bwd(const float* dy, const float* y, float* out, size_t n, int act, void* stream);
int nemo_b200_scale_add(float* y, const float* x, float a, size_t n, void* stream);       /* y += a*x */
int nemo_b200_combine_losses(const float* terms, const float* weights, int n, float* out, void* stream);
/* Gradient rows of the (N+1, 147) MotionNet output beyond the 144 rot6d columns (the fit models subtract the phase-0
 * translation, trans - trans_0, nemo/neural_motion_model.py:3755-3766; row N of the MLP batch is that phase-0 sample):
 *   dY[n, 144:147] = dtrans[n],  dY[n, 147..ldy) = 0  (n < N);   dY[N, 0:144] = 0,  dY[N, 144:147] = -sum_n dtrans[n].
 * dtrans (N,3) may be NULL (= zeros).  Columns 0..143 of rows n < N belong to nemo_b200_rot6d_bwd.
 * scratch: 4 floats the caller has zeroed (column sums + a ticket; left dirty). */
int nemo_b200_dy_tail(const float* dtrans, float* dY, int ldy, int N, float* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEMO_B200_H_ */
