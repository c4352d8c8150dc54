"""torch.autograd.Function wrappers over the C ABI (include/nemo_b200.h).

Every op runs a hand-written sm_100a kernel forward AND backward; there is no PyTorch/CPU
fallback (a CPU tensor raises).  The reference functions each op stands in for are cited per class
(paths under /root/reference).  PyTorch is used for memory, streams and autograd bookkeeping only.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import assets
from ._lib import lib, require_cuda

VP = 6912
NV = 6890
ACT = {None: 0, 'none': 0, 'relu': 1, 'leaky_relu': 2}
RBF_BASES = ['gaussian', 'linear', 'quadratic', 'inverse_quadratic', 'multiquadric', 'inverse_multiquadric',
             'spline', 'poisson_one', 'poisson_two', 'matern32', 'matern52']
KP_LOSSES = ['rmse', 'rmse_resized', 'mse', 'rmse_robust', 'mse_robust', 'mse_robust_resized']


def _f32c(t):
    if t.dtype != torch.float32:
        raise RuntimeError('nemo_b200 ops are fp32 (like the reference); got %s' % t.dtype)
    return t.contiguous()


def _new(shape, like, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device=like.device)


_POSE_BLEND_H16 = bool(int(os.environ.get('NEMO_B200_POSEBLEND_H16', '1')))
A_SCALE_H16 = float(2 ** 13)      # the pose feature R - I lies in [-2, 2]


def pose_blend(F, C, PO, N, ldf):
    """PO (N, 3VP) = F (N, 207; row stride ldf) @ posedirs (lbs.py:229-233) for ROTATION pose features (|R - I| <= 2):
    the fp16-pair tensor-core kernel with the constant operand split once (`SMPLConstants.PDh / PDl`), half the MMAs of the
    3xTF32 GEMM.  NEMO_B200_POSEBLEND_H16=0 keeps the 3xTF32 kernel."""
    if _POSE_BLEND_H16 and C.PDh is not None:
        lib.gemm_h16(F, C.PDh, C.PDl, None, PO, N, 3 * VP, 207, ldf, C.PDh.shape[1], 3 * VP, A_SCALE_H16,
                     1.0 / (A_SCALE_H16 * C.pd_scale))
    else:
        gemm(F, C.PD, PO, N, 3 * VP, 207, ldf, 3 * VP, 3 * VP)


def pose_blend_dgrad(dPO, C, dF, N, ldf, bound):
    """dF (N, 207; row stride ldf) = dPO (N, 3VP) @ posedirs^T for an upstream gradient with a KNOWN bound |dPO| <= bound
    (the vertex L1 of vposer_loss: dv = +-s or 0 and the blended rotation has entries <= 1, so |R^T dv| <= 3 s): the
    fp16-pair kernel with split-K.  Anything else goes through the 3xTF32 GEMM."""
    if _POSE_BLEND_H16 and C.PDh is not None and bound > 0 and np.isfinite(bound):
        a_scale = float(2.0 ** (13 - int(np.ceil(np.log2(bound)))))
        lib.gemm_h16(dPO, C.PDgh, C.PDgl, None, dF, N, 207, 3 * VP, 3 * VP, 3 * VP, ldf, a_scale,
                     1.0 / (a_scale * C.pd_scale))
    else:
        gemm(dPO, C.PD, dF, N, 207, 3 * VP, 3 * VP, 3 * VP, ldf, transB=True)


def gemm(A, B, C, M, N, K, lda, ldb, ldc, transA=False, transB=False, bias=None, act=0, gate=None, ldg=0,
         gate_slope=0.0, accumulate=False):
    """Raw C-ABI GEMM on device pointers/tensors (see nemo_b200_gemm)."""
    lib.gemm(A, B, bias, C, M, N, K, lda, ldb, ldc, int(transA), int(transB), act, gate, ldg, float(gate_slope),
             int(accumulate))


# ----------------------------------------------------------------------------- rotations
class _Rot6d(torch.autograd.Function):
    """hmr.geometry.rot6d_to_rotmat (hmr/geometry.py:47-61)."""

    @staticmethod
    def forward(ctx, x):
        x = _f32c(x.reshape(-1, 6))
        R = _new((x.shape[0], 3, 3), x)
        lib.rot6d_fwd(x, 6, R, 9, 1, x.shape[0])
        ctx.save_for_backward(x)
        return R

    @staticmethod
    def backward(ctx, dR):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        lib.rot6d_bwd(x, 6, _f32c(dR), 9, dx, 6, 1, 0, x.shape[0])
        return dx


def rot6d_to_rotmat(x):
    require_cuda()
    shp = x.shape
    R = _Rot6d.apply(x)
    return R if x.dim() <= 2 and shp[-1] == 6 else R


class _Rotmat2AA(torch.autograd.Function):
    """hmr.geometry.rotation_matrix_to_angle_axis (hmr/geometry.py:181-346; nan_to_zero) and
    human_body_prior.tools.rotation_tools.matrot2aa (tgm_conversion.py:219-376)."""

    @staticmethod
    def forward(ctx, R, nan_to_zero):
        R = _f32c(R.reshape(-1, 3, 3))
        aa = _new((R.shape[0], 3), R)
        lib.rotmat2aa_fwd(R, 9, aa, 3, 1, R.shape[0], int(nan_to_zero))
        ctx.save_for_backward(R)
        ctx.nz = int(nan_to_zero)
        return aa

    @staticmethod
    def backward(ctx, daa):
        (R,) = ctx.saved_tensors
        dR = torch.empty_like(R)
        lib.rotmat2aa_bwd(R, 9, _f32c(daa), 3, dR, 9, 1, 0, R.shape[0], ctx.nz)
        return dR, None


def rotmat_to_aa(R, nan_to_zero=True):
    require_cuda()
    return _Rotmat2AA.apply(R, nan_to_zero)


class _Rodrigues(torch.autograd.Function):
    """kind 0: hmr.geometry.batch_rodrigues (hmr/geometry.py:9-45);
    kind 1: human_body_prior.body_model.lbs.batch_rodrigues (lbs.py:303-334)."""

    @staticmethod
    def forward(ctx, aa, kind):
        aa = _f32c(aa.reshape(-1, 3))
        R = _new((aa.shape[0], 3, 3), aa)
        lib.rodrigues_fwd(aa, 3, R, 9, 1, aa.shape[0], kind)
        ctx.save_for_backward(aa)
        ctx.kind = kind
        return R

    @staticmethod
    def backward(ctx, dR):
        (aa,) = ctx.saved_tensors
        d = torch.empty_like(aa)
        lib.rodrigues_bwd(aa, 3, _f32c(dR), 9, d, 3, 1, 0, aa.shape[0], ctx.kind)
        return d, None


def batch_rodrigues(aa, kind=0):
    require_cuda()
    return _Rodrigues.apply(aa, kind)


# ----------------------------------------------------------------------------- projection
class _Project(torch.autograd.Function):
    """perspective_projection: hmr/geometry.py:78-106 (eps 0), utils/geometry.py:63-93 (eps 1e-8)."""

    @staticmethod
    def forward(ctx, points, rot, trans, focal, centre, eps):
        B, P = points.shape[:2]
        points, rot, trans = _f32c(points), _f32c(rot.reshape(B, 9)), _f32c(trans)
        focal = _f32c(torch.as_tensor(focal, dtype=torch.float32, device=points.device).reshape(-1).expand(B))
        centre = _f32c(centre.expand(B, 2))
        out = _new((B, P, 2), points)
        lib.project_fwd(points, rot, trans, focal, centre, float(eps), out, B, P)
        ctx.save_for_backward(points, rot, trans, focal)
        ctx.eps = float(eps)
        return out

    @staticmethod
    def backward(ctx, dout):
        points, rot, trans, focal = ctx.saved_tensors
        B, P = points.shape[:2]
        dp, dr, dt = torch.empty_like(points), torch.empty_like(rot), torch.empty_like(trans)
        lib.project_bwd(points, rot, trans, focal, ctx.eps, _f32c(dout), dp, dr, dt, B, P)
        return dp, dr.view(B, 3, 3), dt, None, None, None


def perspective_projection(points, rotation, translation, focal_length, camera_center, epsilon=0.0):
    require_cuda()
    return _Project.apply(points, rotation, translation, focal_length, camera_center, epsilon)


# ----------------------------------------------------------------------------- linear layers
class _Linear(torch.autograd.Function):
    """y = act(x W^T + b): nn.Linear (+ReLU / LeakyReLU(0.01)) of FCNN/MotionNet
    (nemo/neural_motion_model.py:58-71,130-148) and VPoser (vposer_model.py:63-88)."""

    @staticmethod
    def forward(ctx, x, W, b, act):
        x, W = _f32c(x), _f32c(W)
        M, K = x.shape
        N = W.shape[0]
        y = _new((M, N), x)
        gemm(x, W, y, M, N, K, K, K, N, transB=True, bias=None if b is None else _f32c(b), act=act)
        ctx.save_for_backward(x, W, y)
        ctx.act, ctx.has_b = act, b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        M, K = x.shape
        N = W.shape[0]
        dy = _f32c(dy)
        if ctx.act:
            g = torch.empty_like(dy)
            lib.act_bwd(dy, y, g, dy.numel(), ctx.act)
            dy = g
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = _new((M, K), x)
            gemm(dy, W, dx, M, K, N, N, K, K)
        if ctx.needs_input_grad[1]:
            dW = _new((N, K), x)
            gemm(dy, x, dW, N, K, M, N, K, K, transA=True)
        if ctx.has_b and ctx.needs_input_grad[2]:
            db = _new((N,), x)
            lib.colsum(dy, db, M, N, N, 0)
        return dx, dW, db, None


def linear(x, W, b=None, act=None):
    require_cuda()
    return _Linear.apply(x, W, b, ACT[act] if not isinstance(act, int) else act)


# ----------------------------------------------------------------------------- RBF / phase warp
class _RBF(torch.autograd.Function):
    """nemo.rbf.RBF.forward (nemo/rbf.py:47-56) with the 11 bases of :62-118."""

    @staticmethod
    def forward(ctx, x, centres, log_sigmas, basis):
        x, centres, log_sigmas = _f32c(x.reshape(-1)), _f32c(centres.reshape(-1)), _f32c(log_sigmas)
        N, D = x.shape[0], log_sigmas.shape[0]
        out = _new((N, D), x)
        lib.rbf_fwd(x, centres, log_sigmas, D, basis, out, N)
        ctx.save_for_backward(x, centres, log_sigmas)
        ctx.basis = basis
        return out

    @staticmethod
    def backward(ctx, dout):
        x, centres, log_sigmas = ctx.saved_tensors
        N, D = x.shape[0], log_sigmas.shape[0]
        dx = torch.empty_like(x)
        dls = torch.zeros_like(log_sigmas)
        lib.rbf_bwd(x, centres, log_sigmas, D, ctx.basis, _f32c(dout), dx, dls, N)
        return dx.view(N, 1), None, dls, None


def rbf(x, centres, log_sigmas, basis):
    require_cuda()
    return _RBF.apply(x, centres, log_sigmas, RBF_BASES.index(basis) if isinstance(basis, str) else basis)


class _Monotonic(torch.autograd.Function):
    """monotonic_network.MonotonicNetwork.forward (monotonic_network.py:23-39)."""

    @staticmethod
    def forward(ctx, x, shifts, scales):
        x, shifts, scales = _f32c(x.reshape(-1)), _f32c(shifts), _f32c(scales)
        out = torch.empty_like(x)
        lib.monotonic_fwd(x, shifts, scales, shifts.shape[0], out, x.shape[0])
        ctx.save_for_backward(x, shifts, scales)
        return out.view(-1, 1)

    @staticmethod
    def backward(ctx, dout):
        x, shifts, scales = ctx.saved_tensors
        dx, dsh, dsc = torch.empty_like(x), torch.zeros_like(shifts), torch.zeros_like(scales)
        scratch = _new((2,), x)
        lib.monotonic_bwd(x, shifts, scales, shifts.shape[0], _f32c(dout.reshape(-1)), dx, dsh, dsc, scratch,
                          x.shape[0])
        return dx.view(-1, 1), dsh, dsc


def monotonic(x, shifts, scales):
    require_cuda()
    return _Monotonic.apply(x, shifts, scales)


# ----------------------------------------------------------------------------- GMM prior
class _GMM(torch.autograd.Function):
    """hmr.smplify.prior.MaxMixturePrior.merged_log_likelihood (hmr/smplify/prior.py:181-196)."""

    @staticmethod
    def forward(ctx, pose, means, precisions, nll_weights):
        pose = _f32c(pose)
        N, D = pose.shape
        M = means.shape[0]
        ll = _new((N,), pose)
        g = _new((N, D), pose)
        lib.gmm_prior(pose, D, _f32c(means), _f32c(precisions), _f32c(nll_weights.reshape(-1)), M, D, ll, 1.0, g, D,
                      0, N)
        ctx.save_for_backward(g)
        return ll

    @staticmethod
    def backward(ctx, dll):
        (g,) = ctx.saved_tensors
        return g * dll[:, None], None, None, None


def gmm_prior(pose, means, precisions, nll_weights):
    require_cuda()
    return _GMM.apply(pose, means, precisions, nll_weights)


# ----------------------------------------------------------------------------- SMPL constants + LBS
class SMPLConstants:
    """Device-resident, kernel-ready SMPL data (HBM layout):

      vt      (3,VP)      v_template, planar, zero padded
      Sp      (10,3VP)    shapedirs as rows of the planar layout
      PD      (207,3VP)   posedirs with columns permuted interleaved(v,c) -> planar(c,v)
      basis   (217,3VP)   [PD; Sp] for per-frame betas
      Jt      (72,)       J_regressor . v_template        JSt (10,72)  J_regressor . shapedirs (transposed)
      Wt      (24,VP)     lbs_weights^T
      readout             joint read-out operators (make_readout)
    """

    def __init__(self, sm: dict, device='cuda'):
        require_cuda()
        dev = torch.device(device)
        f = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32)
        vt = f(sm['v_template'])
        sd = f(sm['shapedirs'])
        pd = f(sm['posedirs'])
        jr = f(sm['J_regressor'])
        w = f(sm['lbs_weights'])
        jx = f(sm['J_regressor_extra'])
        self.num_betas = sd.shape[2]

        def planar(rows):  # (R, 6890, 3) -> (R, 3*VP)
            out = torch.zeros(rows.shape[0], 3, VP)
            out[:, :, :NV] = rows.permute(0, 2, 1)
            return out.reshape(rows.shape[0], 3 * VP)

        self.vt = planar(vt[None])[0].view(3, VP).to(dev).contiguous()
        self.Sp = planar(sd.permute(2, 0, 1)).to(dev).contiguous()
        self.PD = planar(pd.view(207, NV, 3)).to(dev).contiguous()
        self.basis = torch.cat([self.PD, self.Sp], 0).contiguous()
        # posedirs^T as an fp16 PAIR (hi = rn16(x), lo = rn16(x - hi)) of 2^k posedirs, K-major (3VP, 224): the constant
        # operand of the fp16-pair pose-blend GEMM (csrc/gemm_h16.cu); the largest entry lands in [2^13, 2^14)
        pmax = float(self.PD.abs().max())
        self.PDh = self.PDl = None
        self.pd_scale = 1.0
        if pmax > 0 and np.isfinite(pmax):
            self.pd_scale = float(2.0 ** (13 - int(np.floor(np.log2(pmax)))))
            x = torch.zeros(3 * VP, 224, device=dev)
            x[:, :207] = self.PD.t() * self.pd_scale
            self.PDh = x.half().contiguous()
            self.PDl = (x - self.PDh.float()).half().contiguous()
            # the same pair in posedirs' own orientation (207, 3VP): K-major operand of the dgrad (K = 3VP)
            y = self.PD * self.pd_scale
            self.PDgh = y.half().contiguous()
            self.PDgl = (y - self.PDgh.float()).half().contiguous()
        self.Jt = (jr.double() @ vt.double()).float().reshape(72).to(dev)
        self.JSt = torch.einsum('jv,vcl->ljc', jr.double(), sd.double()).float().reshape(self.num_betas, 72).to(dev).contiguous()
        wt = torch.zeros(24, VP)
        wt[:, :NV] = w.t()
        self.Wt = wt.to(dev).contiguous()
        # host float64 copies for the read-out operators: basis (217, 3, NV), template (3, NV), weights (NV, 24)
        self._host = dict(basis=torch.cat([pd.view(207, NV, 3), sd.permute(2, 0, 1)], 0).permute(0, 2, 1).double().contiguous(),
                          vt=vt.t().double().contiguous(), w=w.double())
        self.set_readouts(np.asarray(sm['vertex_joint_ids']), jx)
        self.no_readout = self.make_readout([], torch.zeros(0, NV))
        self.parents = torch.as_tensor(np.asarray(sm['parents']), dtype=torch.int32).to(dev)
        self.joint_map = torch.as_tensor(np.asarray(sm['joint_map']), dtype=torch.int32).to(dev)
        self.faces = np.asarray(sm['faces'])
        self.device = dev
        # fp32 originals for the API-facing buffers of the module mirrors
        self.raw = {k: f(sm[k]).to(dev) for k in ('v_template', 'shapedirs', 'posedirs', 'J_regressor',
                                                   'lbs_weights', 'J_regressor_extra')}

    def make_readout(self, pick_ids, jx_rows):
        """Read-out operators for a set of vertex picks (smplx VertexJointSelector) and regressor rows
        (hmr/smpl.py:33-36), see csrc/readout.cu.  A read-out is linear in the mesh, so it is pre-contracted with
        the pose / shape blend bases and the skinning weights once (float64 on the host):

          MZ (207+nb, ldz), Z0 (ldz,): Z = [pose feature | betas] . MZ + Z0, columns [np x 3 | nx x 24 x 3]
          Wp (np,24) weights of the picked vertices, Sj (nx,24) = sum_v Jx[r,v] W[v,j], rowsum (nx,)
        """
        pick_ids = np.asarray(pick_ids, dtype=np.int64).reshape(-1)
        if len(set(pick_ids.tolist())) != len(pick_ids):
            raise ValueError('vertex pick ids must be unique')
        jx_rows = torch.as_tensor(jx_rows, dtype=torch.float32).reshape(-1, NV).cpu()
        n_p, n_x = len(pick_ids), jx_rows.shape[0]
        H = self._host
        basis, vt, w = H['basis'], H['vt'], H['w']          # (K,3,NV), (3,NV), (NV,24)
        K = basis.shape[0]
        ldz = max(4, (3 * n_p + 72 * n_x + 3) // 4 * 4)
        MZ = torch.zeros(K, ldz, dtype=torch.float64)
        Z0 = torch.zeros(ldz, dtype=torch.float64)
        ids = torch.as_tensor(pick_ids)
        if n_p:
            MZ[:, :3 * n_p] = basis[:, :, ids].permute(0, 2, 1).reshape(K, 3 * n_p)
            Z0[:3 * n_p] = vt[:, ids].t().reshape(-1)
        Sj = torch.zeros(n_x, 24, dtype=torch.float64)
        for r in range(n_x):
            B = jx_rows[r].double()[:, None] * w                                  # (NV,24)
            o = 3 * n_p + 72 * r
            MZ[:, o:o + 72] = (basis.reshape(K * 3, NV) @ B).reshape(K, 3, 24).permute(0, 2, 1).reshape(K, 72)
            Z0[o:o + 72] = (vt @ B).t().reshape(-1)
            Sj[r] = B.sum(0)
        dev = self.vt.device
        f32 = lambda t: t.float().to(dev).contiguous()
        return dict(np=n_p, nx=n_x, ldz=ldz, MZ=f32(MZ), Z0=f32(Z0), Wp=f32(w[ids]) if n_p else None,
                    Sj=f32(Sj) if n_x else None, rowsum=jx_rows.double().sum(1).float().to(dev))

    def set_readouts(self, pick_ids, jx_rows):
        self.readout = self.make_readout(pick_ids, jx_rows)


class _LBS(torch.autograd.Function):
    """lbs (human_body_prior/body_model/lbs.py:164-254) on rotation matrices + the joint read-outs of
    smplx VertexJointSelector / hmr.smpl.SMPL (hmr/smpl.py:33-36).

    rot (N,24,3,3); betas (1,nb) shared or (N,nb) per frame.
    returns vertices (N,6890,3) [or an empty tensor], posed joints JP (N,24,3), read-outs JX (N,np+nx,3).
    The read-outs never touch the mesh (csrc/readout.cu); when no gradient arrives through the vertices the
    backward skips the skinning kernel and the 20736-wide pose-blend dgrad altogether."""

    @staticmethod
    def forward(ctx, rot, betas, C, readout, want_vertices):
        N = rot.shape[0]
        rot = _f32c(rot.reshape(N, 24, 9))
        betas = _f32c(betas)
        per_frame = betas.shape[0] != 1
        nb = betas.shape[1]
        dev = rot.device
        ro = readout
        nq = ro['np'] + ro['nx']
        K = 207 + nb if per_frame else 207
        ldf = (K + 3) // 4 * 4                # 16-byte rows: the GEMM reads F through TMA
        F = torch.zeros((N, ldf), device=dev)
        zb = None
        if per_frame:
            if betas.shape[0] != N:
                raise ValueError('betas batch %d does not match pose batch %d' % (betas.shape[0], N))
            F[:, 207:K] = betas
            basis, base = C.basis, C.vt
            J = _new((N, 72), rot)
            gemm(betas, C.JSt, J, N, 72, nb, nb, 72, 72, bias=C.Jt)
            js = 72
            zb = ro['Z0']
        else:
            basis = C.PD
            base = _new((3, VP), rot)
            gemm(betas, C.Sp, base, 1, 3 * VP, nb, nb, 3 * VP, 3 * VP, bias=C.vt.view(-1))
            J = _new((72,), rot)
            gemm(betas, C.JSt, J, 1, 72, nb, nb, 72, 72, bias=C.Jt)
            js = 0
            if nq:
                zb = _new((ro['ldz'],), rot)
                gemm(betas, ro['MZ'][207:], zb, 1, ro['ldz'], nb, nb, ro['ldz'], ro['ldz'], bias=ro['Z0'])
        A = _new((N, 24, 12), rot)
        JP = _new((N, 24, 3), rot)
        lib.pose_chain_fwd(rot, None, 0, 0, None, J, js, C.parents, None, F, ldf, A, JP, N)   # pose feature + chain
        PO = None
        V = torch.empty(0, device=dev)
        if want_vertices:
            PO = _new((N, 3 * VP), rot)
            gemm(F, basis, PO, N, 3 * VP, K, ldf, 3 * VP, 3 * VP)
            V = _new((N, NV, 3), rot)
            lib.skin_fwd(A, PO, base, C.Wt, V, N)
        JX = _new((N, nq, 3), rot)
        Z = None
        if nq:
            Z = _new((N, ro['ldz']), rot)
            gemm(F, ro['MZ'], Z, N, ro['ldz'], K, ldf, ro['ldz'], ro['ldz'], bias=zb)
            lib.readout_fwd(A, Z, ro['ldz'], ro['Wp'], ro['np'], ro['Sj'], ro['nx'], JX, N)
        ctx.save_for_backward(rot, betas, PO, base, J, A, Z)
        ctx.C, ctx.ro, ctx.per_frame, ctx.js, ctx.K = C, ro, per_frame, js, K
        return V, JP, JX

    @staticmethod
    def backward(ctx, dV, dJP, dJX):
        rot, betas, PO, base, J, A, Z = ctx.saved_tensors
        C, ro, K = ctx.C, ctx.ro, ctx.K
        N, nb = rot.shape[0], betas.shape[1]
        ldz = ro['ldz']
        dV = None if dV is None or dV.numel() == 0 or PO is None else _f32c(dV)
        dJX = None if dJX is None or dJX.numel() == 0 or Z is None else _f32c(dJX)
        dJP = None if dJP is None else _f32c(dJP)
        dA = dZ = dPO = None
        if dJX is not None:
            dA, dZ = _new((N, 24, 12), rot), _new((N, ldz), rot)
            lib.readout_bwd(A, Z, ldz, ro['Wp'], ro['np'], ro['Sj'], ro['nx'], dJX, dA, 0, dZ, N)
        if dV is not None:
            dPO, dAs = _new((N, 3 * VP), rot), _new((N, 24, 12), rot)
            lib.skin_bwd(A, PO, base, C.Wt, dV, None, 0.0, dPO, dAs, N)
            dA = dAs if dA is None else dA.add_(dAs)
        dR = _new((N, 24, 9), rot)
        dJ = _new((N, 72), rot) if ctx.per_frame else torch.zeros(72, device=rot.device)
        # d(pose feature) into its own aligned rows (tensor-core GEMMs with the TMA epilogue), added into dR[:, 1:] by the
        # chain backward as it writes dR
        dF = None
        if dPO is not None or dZ is not None:
            dF = _new((N, 208), rot)
            if dPO is not None:
                gemm(dPO, C.PD, dF, N, 207, 3 * VP, 3 * VP, 3 * VP, 208, transB=True)
            if dZ is not None:
                gemm(dZ, ro['MZ'], dF, N, 207, ldz, ldz, ldz, 208, transB=True, accumulate=dPO is not None)
        lib.chain_bwd2(rot, J, ctx.js, C.parents, dA, None, dJP, dF, 208, dR, dJ, N)
        dbetas = None
        if ctx.needs_input_grad[1]:
            MZs = ro['MZ'][207:]
            if ctx.per_frame:
                dbetas = _new((N, nb), rot)
                gemm(dJ, C.JSt, dbetas, N, nb, 72, 72, 72, nb, transB=True)
                if dPO is not None:
                    gemm(dPO, C.Sp, dbetas, N, nb, 3 * VP, 3 * VP, 3 * VP, nb, transB=True, accumulate=True)
                if dZ is not None:
                    gemm(dZ, MZs, dbetas, N, nb, ldz, ldz, ldz, nb, transB=True, accumulate=True)
            else:
                dbetas = _new((1, nb), rot)
                gemm(dJ, C.JSt, dbetas, 1, nb, 72, 72, 72, nb, transB=True)
                if dPO is not None:
                    dbase = _new((3 * VP,), rot)
                    lib.colsum(dPO, dbase, N, 3 * VP, 3 * VP, 0)
                    gemm(dbase, C.Sp, dbetas, 1, nb, 3 * VP, 3 * VP, 3 * VP, nb, transB=True, accumulate=True)
                if dZ is not None:
                    dzb = _new((ldz,), rot)
                    lib.colsum(dZ, dzb, N, ldz, ldz, 0)
                    gemm(dzb, MZs, dbetas, 1, nb, ldz, ldz, ldz, nb, transB=True, accumulate=True)
        return dR.view(N, 24, 3, 3), dbetas, None, None, None


def lbs_rotmats(rot, betas, C: SMPLConstants, readout=None, want_vertices=True):
    require_cuda()
    V, JP, JX = _LBS.apply(rot, betas, C, readout or C.readout, want_vertices)
    return (V if want_vertices else None), JP, JX


class _JointsGather(torch.autograd.Function):
    """joints = cat(JP, JX)[:, joint_map] (+ transl): smplx SMPL.forward + hmr/smpl.py:33-37."""

    @staticmethod
    def forward(ctx, JP, JX, jmap, np_, transl, rowsum):
        N = JP.shape[0]
        JP, JX = _f32c(JP), _f32c(JX)
        nout = jmap.shape[0]
        out = _new((N, nout, 3), JP)
        tr = None if transl is None else _f32c(transl)
        lib.joints_gather_fwd(JP, JX if JX.numel() else None, JX.shape[1], np_, jmap, nout, tr, rowsum, out, N)
        ctx.meta = (JX.shape[1], np_, nout, transl is not None)
        ctx.save_for_backward(jmap, rowsum if rowsum is not None else torch.empty(0, device=JP.device))
        return out

    @staticmethod
    def backward(ctx, dout):
        jmap, rowsum = ctx.saved_tensors
        nx_raw, np_, nout, has_t = ctx.meta
        N = dout.shape[0]
        dout = _f32c(dout)
        dJP, dJX = _new((N, 24, 3), dout), _new((N, nx_raw, 3), dout)
        dtr = _new((N, 3), dout) if has_t else None
        lib.joints_gather_bwd(dout, nx_raw, np_, jmap, nout, rowsum if rowsum.numel() else None, dJP,
                              dJX if nx_raw else None, dtr, N)
        return dJP, dJX, None, None, dtr, None


def joints_gather(JP, JX, jmap, np_, transl=None, rowsum=None):
    return _JointsGather.apply(JP, JX, jmap, np_, transl, rowsum)
