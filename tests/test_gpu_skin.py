"""K7 (tcgen05 skinning forward / fused L1 / backward) and the mesh-free joint read-outs, called through the C ABI on
ragged batch sizes and checked against a float64 restatement of lbs.py:236-252 / hmr/smpl.py:33-36.  Needs a B200."""
import numpy as np
import pytest
import torch

from nemo_cvpr2023_b200 import assets
from tests.util import TOL_GRAD_REL, TOL_VERT_ABS, assert_abs, assert_grad

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def ctx():
    import nemo_cvpr2023_b200.functional as Fn
    from nemo_cvpr2023_b200._lib import lib
    sm = assets.make_smpl_assets(0)
    return Fn, lib, Fn.SMPLConstants(sm, DEV), sm


def _inputs(N, seed):
    r = np.random.RandomState(seed)
    A = r.standard_normal((N, 24, 12)) * 0.5
    PO = np.zeros((N, 3, 6912))
    PO[:, :, :6890] = 0.01 * r.standard_normal((N, 3, 6890))
    return A, PO


def _ref_skin(A, PO, sm):
    """float64: v[n,v] = sum_j W[v,j] (A[n,j]_3x3 (vt[v] + PO[n,:,v]) + A[n,j]_t)"""
    W = np.asarray(sm['lbs_weights'], np.float64)
    vt = np.asarray(sm['v_template'], np.float64)
    p = vt[None] + PO[:, :, :6890].transpose(0, 2, 1)                # (N,V,3)
    T = np.einsum('vj,njk->nvk', W, A.reshape(len(A), 24, 12)).reshape(len(A), -1, 3, 4)
    return np.einsum('nvrc,nvc->nvr', T[..., :3], p) + T[..., 3], p, T


def g32(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32, device=DEV).contiguous()


@pytest.mark.parametrize('N', [1, 7, 8, 37, 300])
def test_skin_forward_and_backward_ragged(ctx, N):
    Fn, lib, C, sm = ctx
    A, PO = _inputs(N, N)
    A32, PO32 = g32(A), g32(PO.reshape(N, -1))
    A, PO = A32.double().cpu().numpy(), PO32.double().cpu().numpy().reshape(N, 3, 6912)
    V = torch.full((N, 6890, 3), float('nan'), device=DEV)
    lib.skin_fwd(A32, PO32, C.vt, C.Wt, V, N)
    ref, p, T = _ref_skin(A, PO, sm)
    assert_abs(V.double().cpu().numpy(), ref, TOL_VERT_ABS, 'vertices N=%d' % N)
    # backward: dPO = R^T dv (planar), dA[n][j][r*4+c] = sum_v W[v,j] dv[n,v,r] [p;1][c]
    r = np.random.RandomState(100 + N)
    dV32 = g32(r.standard_normal((N, 6890, 3)))
    dV = dV32.double().cpu().numpy()
    dPO = torch.full((N, 3 * 6912), float('nan'), device=DEV)
    dAs = torch.full((N, 24, 12), float('nan'), device=DEV)
    lib.skin_bwd(A32, PO32, C.vt, C.Wt, dV32, None, 0.0, dPO, dAs, N)
    dPO_ref = np.einsum('nvrc,nvr->ncv', T[..., :3], dV)
    got = dPO.double().cpu().numpy().reshape(N, 3, 6912)
    assert_grad(got[:, :, :6890], dPO_ref, TOL_GRAD_REL, 'dPO N=%d' % N)
    assert np.all(got[:, :, 6890:] == 0), 'pad vertices must get a zero gradient'
    ph = np.concatenate([p, np.ones_like(p[..., :1])], -1)
    W = np.asarray(sm['lbs_weights'], np.float64)
    dA_ref = np.einsum('vj,nvr,nvc->njrc', W, dV, ph).reshape(N, 24, 12)
    assert_grad(dAs.double().cpu().numpy(), dA_ref, TOL_GRAD_REL, 'dA N=%d' % N)


def test_real_asset_shaped_sparse_weights():
    """the licensed SMPL files have <= 4 non-zero skinning weights per vertex and sparse joint
    regressors.  The tensor-core kernels take the dense (V, 24) matrix whatever its pattern (they are bound by bytes
    and the shared-memory pipe, not by the 24-wide contraction, so a dedicated sparse path would not be faster): the
    same parity contract must hold on such weights -- skinning forward / backward and the mesh-free read-outs."""
    import nemo_cvpr2023_b200.functional as Fn
    from nemo_cvpr2023_b200._lib import lib
    sm = assets.make_smpl_assets(3, sparse=True)
    assert (np.asarray(sm['lbs_weights']) > 0).sum(1).max() <= 4
    C = Fn.SMPLConstants(sm, DEV)
    N = 45
    A, PO = _inputs(N, 77)
    A32, PO32 = g32(A), g32(PO.reshape(N, -1))
    A, PO = A32.double().cpu().numpy(), PO32.double().cpu().numpy().reshape(N, 3, 6912)
    V = torch.full((N, 6890, 3), float('nan'), device=DEV)
    lib.skin_fwd(A32, PO32, C.vt, C.Wt, V, N)
    ref, p, T = _ref_skin(A, PO, sm)
    assert_abs(V.double().cpu().numpy(), ref, TOL_VERT_ABS, 'vertices (sparse W)')
    r = np.random.RandomState(5)
    dV32 = g32(r.standard_normal((N, 6890, 3)))
    dV = dV32.double().cpu().numpy()
    dPO = torch.full((N, 3 * 6912), float('nan'), device=DEV)
    dAs = torch.full((N, 24, 12), float('nan'), device=DEV)
    lib.skin_bwd(A32, PO32, C.vt, C.Wt, dV32, None, 0.0, dPO, dAs, N)
    dPO_ref = np.einsum('nvrc,nvr->ncv', T[..., :3], dV)
    assert_grad(dPO.double().cpu().numpy().reshape(N, 3, 6912)[:, :, :6890], dPO_ref, TOL_GRAD_REL, 'dPO (sparse W)')
    ph = np.concatenate([p, np.ones_like(p[..., :1])], -1)
    dA_ref = np.einsum('vj,nvr,nvc->njrc', np.asarray(sm['lbs_weights'], np.float64), dV, ph).reshape(N, 24, 12)
    assert_grad(dAs.double().cpu().numpy(), dA_ref, TOL_GRAD_REL, 'dA (sparse W)')
    # whole SMPL forward + backward (shape blend, sparse joint regressors, chain, skinning, read-outs) against the oracle
    from nemo_cvpr2023_b200.hmr.smpl import SMPL
    from oracle import nemo_oracle as O
    model = SMPL(sm, device=DEV)
    smt = O.to_torch(sm)
    n = 13
    rot = O.batch_rodrigues_quat(torch.tensor(0.4 * r.standard_normal((n * 24, 3)), dtype=torch.float32)).view(n, 24, 3, 3)
    betas = torch.tensor(r.standard_normal((1, 10)), dtype=torch.float32)
    tr = torch.tensor(r.standard_normal((n, 3)), dtype=torch.float32)
    rr = rot.clone().requires_grad_(True)
    ref = O.smpl_forward(smt, betas=betas, body_pose=rr[:, 1:], global_orient=rr[:, :1], transl=tr, pose2rot=False)
    wj = torch.tensor(r.standard_normal((n, 49, 3)), dtype=torch.float32)
    wv = torch.tensor(r.standard_normal((n, 6890, 3)), dtype=torch.float32)
    ((ref.joints * wj).sum() + (ref.vertices * wv).sum()).backward()
    rg = rot.to(DEV).requires_grad_(True)
    out = model(betas=betas.to(DEV), body_pose=rg[:, 1:], global_orient=rg[:, :1], transl=tr.to(DEV), pose2rot=False)
    assert_abs(out.vertices.detach().double().cpu().numpy(), ref.vertices.detach().double().numpy(), TOL_VERT_ABS, 'SMPL vertices (sparse assets)')
    assert_abs(out.joints.detach().double().cpu().numpy(), ref.joints.detach().double().numpy(), TOL_VERT_ABS, 'SMPL joints (sparse assets)')
    ((out.joints * wj.to(DEV)).sum() + (out.vertices * wv.to(DEV)).sum()).backward()
    assert_grad(rg.grad.double().cpu().numpy(), rr.grad.double().numpy(), TOL_GRAD_REL, 'd rotmats (sparse assets)')


@pytest.mark.parametrize('N', [19, 64])
def test_skin_backward_wide_dynamic_range(ctx, N):
    """The adjoint GEMM runs on scaled fp16 operand pairs: upstream gradients whose magnitude differs by many orders
    between frames (and between the vertices of one frame), and all-zero frames, must come out at fp32 accuracy."""
    Fn, lib, C, sm = ctx
    A, PO = _inputs(N, 40 + N)
    A32, PO32 = g32(A), g32(PO.reshape(N, -1))
    A, PO = A32.double().cpu().numpy(), PO32.double().cpu().numpy().reshape(N, 3, 6912)
    r = np.random.RandomState(7 + N)
    dV = r.standard_normal((N, 6890, 3))
    dV *= 10.0 ** r.uniform(-12, 6, size=(N, 1, 1))                    # per-frame magnitude
    dV *= 10.0 ** (4 * (r.uniform(size=(N, 6890, 1)) > 0.999))        # a few vertices 1e4 x larger
    dV[3] = 0.0
    dV[5, :, 1:] = 0.0
    dV32 = g32(dV)
    dV = dV32.double().cpu().numpy()
    dPO = torch.full((N, 3 * 6912), float('nan'), device=DEV)
    dAs = torch.full((N, 24, 12), float('nan'), device=DEV)
    lib.skin_bwd(A32, PO32, C.vt, C.Wt, dV32, None, 0.0, dPO, dAs, N)
    _, p, T = _ref_skin(A, PO, sm)
    ph = np.concatenate([p, np.ones_like(p[..., :1])], -1)
    W = np.asarray(sm['lbs_weights'], np.float64)
    dA_ref = np.einsum('vj,nvr,nvc->njrc', W, dV, ph).reshape(N, 24, 12)
    got = dAs.double().cpu().numpy()
    assert np.all(got[3] == 0)
    for n in range(N):                                                  # per frame: the frames differ by 1e18
        if np.abs(dA_ref[n]).max() > 0:
            assert_grad(got[n], dA_ref[n], TOL_GRAD_REL, 'dA frame %d' % n)
    dPO_ref = np.einsum('nvrc,nvr->ncv', T[..., :3], dV)
    gp = dPO.double().cpu().numpy().reshape(N, 3, 6912)[:, :, :6890]
    for n in range(N):
        if np.abs(dPO_ref[n]).max() > 0:
            assert_grad(gp[n], dPO_ref[n], TOL_GRAD_REL, 'dPO frame %d' % n)


@pytest.mark.parametrize('N', [1, 5, 36, 131])
def test_fused_l1_and_sign_backward(ctx, N):
    Fn, lib, C, sm = ctx
    A, PO = _inputs(N, 10 + N)
    A2, PO2 = _inputs(N, 20 + N)
    a32, p32, b32, q32 = g32(A), g32(PO.reshape(N, -1)), g32(A2), g32(PO2.reshape(N, -1))
    A, A2 = a32.double().cpu().numpy(), b32.double().cpu().numpy()
    PO, PO2 = p32.double().cpu().numpy().reshape(N, 3, 6912), q32.double().cpu().numpy().reshape(N, 3, 6912)
    signs = torch.zeros(N, 3, 2, 6912 // 32, dtype=torch.int32, device=DEV)
    l1 = torch.zeros(1, device=DEV)
    lib.skin_l1_fwd(a32, p32, b32, q32, C.vt, C.Wt, signs, l1, N)
    va, p, T = _ref_skin(A, PO, sm)
    vb, _, _ = _ref_skin(A2, PO2, sm)
    ref = np.abs(va - vb).sum()
    assert abs(float(l1) - ref) <= 1e-5 * ref, (float(l1), ref)
    scale = {1: 0.37, 5: 3.0e-7, 36: 2500.0, 131: 0.37}[N]   # the signs-only backward scales its fp16 operands statically
    dPO = torch.empty(N, 3 * 6912, device=DEV)
    dAs = torch.empty(N, 24, 12, device=DEV)
    lib.skin_bwd(a32, p32, C.vt, C.Wt, None, signs, scale, dPO, dAs, N)
    dV = scale * np.sign(va - vb)
    dPO_ref = np.einsum('nvrc,nvr->ncv', T[..., :3], dV)
    assert_grad(dPO.double().cpu().numpy().reshape(N, 3, 6912)[:, :, :6890], dPO_ref, TOL_GRAD_REL, 'dPO (signs)')
    ph = np.concatenate([p, np.ones_like(p[..., :1])], -1)
    dA_ref = np.einsum('vj,nvr,nvc->njrc', np.asarray(sm['lbs_weights'], np.float64), dV, ph).reshape(N, 24, 12)
    assert_grad(dAs.double().cpu().numpy(), dA_ref, TOL_GRAD_REL, 'dA (signs)')


@pytest.mark.parametrize('N', [1, 33])
def test_readouts_match_the_mesh(ctx, N):
    """picked vertices and J_regressor_extra joints computed without the mesh == the same read off the float64 mesh,
    forward and backward (through A and the pose feature)."""
    Fn, lib, C, sm = ctx
    r = np.random.RandomState(7 + N)
    ro = C.readout
    nq, ldz = ro['np'] + ro['nx'], ro['ldz']
    A32 = g32(r.standard_normal((N, 24, 12)) * 0.5)
    F32 = torch.zeros(N, 208, device=DEV)
    F32[:, :207] = g32(r.standard_normal((N, 207)) * 0.3)
    A, F = A32.double().cpu().numpy(), F32.double().cpu().numpy()[:, :207]
    Z = torch.empty(N, ldz, device=DEV)
    Fn.gemm(F32, ro['MZ'], Z, N, ldz, 207, 208, ldz, ldz, bias=ro['Z0'])
    JX = torch.empty(N, nq, 3, device=DEV)
    lib.readout_fwd(A32, Z, ldz, ro['Wp'], ro['np'], ro['Sj'], ro['nx'], JX, N)
    pd = np.asarray(sm['posedirs'], np.float64).reshape(207, 6890, 3)
    PO = np.zeros((N, 3, 6912))
    PO[:, :, :6890] = np.einsum('nk,kvc->ncv', F, pd)
    V, p, T = _ref_skin(A, PO, sm)
    ids = np.asarray(sm['vertex_joint_ids'])
    jx = np.asarray(sm['J_regressor_extra'], np.float64)
    ref = np.concatenate([V[:, ids], np.einsum('rv,nvc->nrc', jx, V)], 1)
    assert_abs(JX.double().cpu().numpy(), ref, TOL_VERT_ABS, 'read-outs')
    # backward against the float64 mesh path
    g = g32(r.standard_normal((N, nq, 3)))
    gd = g.double().cpu().numpy()
    dA, dZ = torch.empty(N, 24, 12, device=DEV), torch.empty(N, ldz, device=DEV)
    lib.readout_bwd(A32, Z, ldz, ro['Wp'], ro['np'], ro['Sj'], ro['nx'], g, dA, 0, dZ, N)
    dF = torch.zeros(N, 207, device=DEV)
    Fn.gemm(dZ, ro['MZ'], dF, N, 207, ldz, ldz, ldz, 207, transB=True)
    dV = np.zeros((N, 6890, 3))
    np.add.at(dV, (slice(None), ids), gd[:, :len(ids)])
    dV += np.einsum('rv,nrc->nvc', jx, gd[:, len(ids):])
    ph = np.concatenate([p, np.ones_like(p[..., :1])], -1)
    dA_ref = np.einsum('vj,nvr,nvc->njrc', np.asarray(sm['lbs_weights'], np.float64), dV, ph).reshape(N, 24, 12)
    dPO_ref = np.einsum('nvrc,nvr->nvc', T[..., :3], dV)
    dF_ref = np.einsum('nvc,kvc->nk', dPO_ref, pd)
    assert_grad(dA.double().cpu().numpy(), dA_ref, TOL_GRAD_REL, 'dA (read-outs)')
    assert_grad(dF.double().cpu().numpy(), dF_ref, TOL_GRAD_REL, 'd pose feature (read-outs)')
