"""Parity of the CUDA path (through the C ABI) against the golden fixtures produced by the unmodified
reference and against the CPU oracle on seeded inputs.  Needs a GPU (B200)."""
import numpy as np
import pytest
import torch

from nemo_cvpr2023_b200 import assets
from oracle import nemo_oracle as O
from oracle.make_golden import VSUB, step_case
from tests.util import (assert_adam_param, T, TOL_GRAD_REL, TOL_LOSS_REL, TOL_VERT_ABS, assert_abs, assert_grad, assert_rel, gold, npy)

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def G(a, grad=False):
    return T(a, grad, device=DEV)


@pytest.fixture(scope='module')
def Fn():
    import nemo_cvpr2023_b200.functional as Fn
    from nemo_cvpr2023_b200._lib import lib
    arch = lib.load().nemo_b200_sm_arch()
    assert arch >= 100, 'expected sm_100 (B200), got %d' % arch
    return Fn


@pytest.fixture(scope='module')
def smpl():
    from nemo_cvpr2023_b200.hmr.smpl import SMPL
    return SMPL('synthetic')


def test_library_rejects_cpu_tensors(Fn):
    with pytest.raises(RuntimeError):
        Fn.rot6d_to_rotmat(torch.zeros(2, 6))


def test_rot6d(Fn):
    g = gold('geometry')
    x = G(g['rot6d_in'], True)
    r = Fn.rot6d_to_rotmat(x)
    (r * G(g['rot6d_w'])).sum().backward()
    assert_abs(npy(r), g['rot6d_out'], 1e-6)
    assert_grad(npy(x.grad), g['rot6d_grad'])


@pytest.mark.parametrize('kind', [0, 1])
def test_rodrigues(Fn, kind):
    g = gold('geometry')
    key = 'rodq' if kind == 0 else 'rods'
    x = G(g['aa_in'], True)
    r = Fn.batch_rodrigues(x, kind)
    (r * G(g['rod_w'])).sum().backward()
    assert_abs(npy(r), g[key + '_out'], 1e-6)
    assert_grad(npy(x.grad), g[key + '_grad'])


@pytest.mark.parametrize('nz', [True, False])
def test_rotmat_to_aa(Fn, nz):
    g = gold('geometry')
    key = 'r2aa_hmr' if nz else 'r2aa_tgm'
    x = G(g['r2aa_in'], True)
    aa = Fn.rotmat_to_aa(x, nz)
    (aa * G(g['r2aa_w'])).sum().backward()
    assert_abs(npy(aa), g[key], 2e-6)
    assert_grad(npy(x.grad), g[key + '_grad'])
    eye = torch.eye(3, device=DEV)[None].requires_grad_(True)
    a0 = Fn.rotmat_to_aa(eye, nz)
    a0.sum().backward()
    assert torch.all(a0 == 0) and torch.isfinite(eye.grad).all()


def test_projection(Fn):
    from nemo_cvpr2023_b200.hmr import geometry as hg
    from nemo_cvpr2023_b200.utils import geometry as ug
    g = gold('geometry')
    p, r, t = G(g['pp_pts'], True), G(g['pp_rot'], True), G(g['pp_tr'], True)
    o = hg.perspective_projection(p, r, t, G(g['pp_fl']), G(g['pp_cc']))
    (o * G(g['pp_w'])).sum().backward()
    assert_abs(npy(o), g['pp_hmr'], 2e-3)
    assert_grad(npy(p.grad), g['pp_hmr_gpts'])
    assert_grad(npy(r.grad), g['pp_hmr_grot'])
    assert_grad(npy(t.grad), g['pp_hmr_gtr'])
    o2 = ug.perspective_projection(G(g['pp_pts']), G(g['pp_rot']), G(g['pp_tr']), G(g['pp_fl']), G(g['pp_cc']))
    assert_abs(npy(o2), g['pp_utils'], 2e-3)
    o3 = ug.perspective_projection_with_K(G(g['pp_pts']), G(g['pp_rot']), G(g['pp_tr']), G(g['pp_K']))
    assert_abs(npy(o3), g['pp_withK'], 2e-3)


@pytest.mark.parametrize('ta,tb', [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_matches_fp64(Fn, ta, tb):
    r = np.random.RandomState(5)
    for (M, N, K) in [(1, 20736, 10), (77, 147, 1000), (300, 207, 6912), (513, 1000, 105), (147, 1000, 513)]:
        A = r.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
        B = r.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
        bias = r.standard_normal(N).astype(np.float32)
        ref = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64) + bias
        ref = np.where(ref > 0, ref, 0.01 * ref)
        C = torch.empty(M, N, device=DEV)
        Fn.gemm(G(A), G(B), C, M, N, K, A.shape[1], B.shape[1], N, bool(ta), bool(tb), bias=G(bias), act=2)
        # fp32 accumulation over K terms of magnitude ~1: a few ulp of the largest output
        assert_abs(npy(C), ref, 2e-5 * np.abs(ref).max(), 'gemm %s' % ((M, N, K),))


def test_lbs_axis_angle(Fn, smpl):
    from nemo_cvpr2023_b200.human_body_prior.body_model import lbs as L
    g = gold('lbs')
    b, p = G(g['betas'], True), G(g['pose'], True)
    v, j = L.lbs(b, p, smpl.v_template, smpl.shapedirs, smpl.posedirs, smpl.J_regressor, smpl.parents,
                 smpl.lbs_weights, pose2rot=True)
    ((v * G(g['dV'])).sum() + (j * G(g['dJ24'])).sum()).backward()
    assert_abs(npy(v)[:, VSUB], g['lbs_v_sub'], TOL_VERT_ABS, 'verts')
    assert_abs(npy(v).sum((1, 2)), g['lbs_v_sum'], 2e-3, 'vertex checksum')
    assert_abs(npy(j), g['lbs_j'], TOL_VERT_ABS, 'joints')
    assert_grad(npy(b.grad), g['lbs_gbetas'], what='d betas')
    assert_grad(npy(p.grad), g['lbs_gpose'], what='d pose')


def test_body_model(Fn):
    from nemo_cvpr2023_b200.human_body_prior.body_model.body_model import BodyModel
    g = gold('lbs')
    bm = BodyModel('synthetic')
    out = bm(root_orient=G(g['pose'][:, :3]), pose_body=G(g['pose'][:, 3:66]), pose_hand=G(g['pose'][:, 66:]),
             betas=G(g['betas']), trans=G(g['transl']))
    assert_abs(npy(out.v)[:, VSUB], g['lbs_v_sub'] + g['transl'][:, None], TOL_VERT_ABS)
    assert_abs(npy(out.Jtr), g['lbs_j'] + g['transl'][:, None], TOL_VERT_ABS)


def test_smpl_forward(Fn, smpl):
    g = gold('lbs')
    b, r = G(g['betas1'], True), G(g['rotmats'], True)
    o = smpl(betas=b, body_pose=r[:, 1:], global_orient=r[:, :1], pose2rot=False)
    ((o.vertices * G(g['dV'])).sum() + (o.joints * G(g['dJ49'])).sum()).backward()
    assert_abs(npy(o.vertices)[:, VSUB], g['smpl_v_sub'], TOL_VERT_ABS)
    assert_abs(npy(o.joints), g['smpl_j49'], TOL_VERT_ABS)
    assert_grad(npy(b.grad), g['smpl_gbetas'], what='d betas')
    assert_grad(npy(r.grad), g['smpl_grot'], what='d rotmats')
    p = G(g['pose'], True)
    o = smpl(betas=G(g['betas']), body_pose=p[:, 3:], global_orient=p[:, :3], transl=G(g['transl']), pose2rot=True)
    ((o.vertices * G(g['dV'])).sum() + (o.joints * G(g['dJ49'])).sum()).backward()
    assert_abs(npy(o.vertices)[:, VSUB], g['smplaa_v_sub'], TOL_VERT_ABS)
    assert_abs(npy(o.joints), g['smplaa_j49'], TOL_VERT_ABS)
    assert_grad(npy(p.grad), g['smplaa_gpose'], what='d pose')
    o = smpl()
    assert_abs(npy(o.vertices)[:, VSUB], g['rest_v_sub'], TOL_VERT_ABS)
    assert_abs(npy(o.joints), g['rest_j49'], TOL_VERT_ABS)
    assert_abs(npy(o.vertices)[0], npy(smpl.v_template), 2e-6)


def test_smpl_ragged_batches_vs_oracle(Fn, smpl):
    """batch sizes that do not fill a frame group / tile (1, 5, 13) against the CPU oracle."""
    sm = O.to_torch(assets.make_smpl_assets(0))
    r = np.random.RandomState(11)
    for n in (1, 5, 13):
        rot = O.batch_rodrigues_quat(torch.tensor(0.4 * r.standard_normal((n * 24, 3)), dtype=torch.float32)).view(n, 24, 3, 3)
        betas = torch.tensor(r.standard_normal((1, 10)), dtype=torch.float32)
        ref = O.smpl_forward(sm, betas=betas, body_pose=rot[:, 1:], global_orient=rot[:, :1], pose2rot=False)
        out = smpl(betas=betas.to(DEV), body_pose=rot[:, 1:].to(DEV), global_orient=rot[:, :1].to(DEV), pose2rot=False)
        assert_abs(npy(out.vertices), npy(ref.vertices), TOL_VERT_ABS, 'n=%d' % n)
        assert_abs(npy(out.joints), npy(ref.joints), TOL_VERT_ABS, 'n=%d' % n)
    out = smpl(betas=torch.zeros(0, 10, device=DEV), body_pose=torch.zeros(0, 23, 3, 3, device=DEV),
               global_orient=torch.zeros(0, 1, 3, 3, device=DEV), pose2rot=False)
    assert out.vertices.shape == (0, 6890, 3) and out.joints.shape == (0, 49, 3)


def test_rbf_and_monotonic(Fn):
    g = gold('modules')
    c = torch.linspace(0, 1, 12, device=DEV)[:, None]
    for nm in Fn.RBF_BASES:
        x, ls = G(g['rbf_x'], True), G(g['rbf_ls'], True)
        y = Fn.rbf(x, c, ls, nm)
        y.sum().backward()
        assert_abs(npy(y), g['rbf_' + nm], 2e-6, nm)
        assert_grad(npy(x.grad), g['rbf_%s_gx' % nm], what=nm)
        assert_grad(npy(ls.grad), g['rbf_%s_gls' % nm], what=nm)
    x, sh, sc = G(g['rbf_x'], True), G(g['mono_sh'], True), G(g['mono_sc'], True)
    y = Fn.monotonic(x, sh, sc)
    (y * G(g['mono_w'])).sum().backward()
    assert_abs(npy(y), g['mono_y'], 2e-6)
    assert_grad(npy(x.grad), g['mono_gx'])
    assert_grad(npy(sh.grad), g['mono_gsh'])
    assert_grad(npy(sc.grad), g['mono_gsc'])


def test_vposer_and_gmm(Fn):
    from nemo_cvpr2023_b200.human_body_prior.models.vposer_model import VPoser
    from nemo_cvpr2023_b200.hmr.smplify.prior import MaxMixturePrior
    g = gold('modules')
    vp = VPoser(state_dict=assets.make_vposer_weights(2))
    x = G(g['vp_in'], True)
    q = vp.encode(x)
    dec = vp.decode(q.mean)
    kl = torch.mean(torch.sum(torch.distributions.kl.kl_divergence(
        q, torch.distributions.normal.Normal(torch.zeros_like(q.mean), torch.ones_like(q.mean))), dim=[1]))
    ((dec['pose_body'] * G(g['vp_w'])).sum() + kl).backward()
    assert_abs(npy(q.mean), g['vp_mu'], 3e-6)
    assert_abs(npy(q.scale), g['vp_sigma'], 3e-6)
    assert_abs(npy(dec['pose_body']), g['vp_dec'], 5e-6)
    assert_rel(float(kl), g['vp_kl'], TOL_LOSS_REL)
    assert_grad(npy(x.grad), g['vp_gin'])
    pr = MaxMixturePrior(num_gaussians=8, gmm=assets.make_gmm(3))
    x = G(g['gmm_in'], True)
    ll = pr(x, None)
    ll.sum().backward()
    assert_abs(npy(ll) / g['gmm_ll'], np.ones_like(g['gmm_ll']), TOL_LOSS_REL)
    assert_grad(npy(x.grad), g['gmm_g'])


def build_model(version):
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence
    args, params, data, view_idx, frame_idx = step_case(version)
    cls = {2: nmm.NemoV2, 3: nmm.NemoV3, 4: nmm.NemoV4}[version]
    m = cls(args, SyntheticMultiViewSequence(data=data), DEV, init_params=params)
    m.engine.compute_betas_grad = True
    return m, torch.tensor(view_idx), torch.tensor(frame_idx)


@pytest.mark.parametrize('version', [2, 3, 4])
def test_step_matches_reference(version):
    g = gold('step_v%d' % version)
    m, vi, fi = build_model(version)
    pred = m.get_preds_batch(vi, fi)
    assert_abs(npy(pred['v'])[:, VSUB], g['pred_v_sub'], TOL_VERT_ABS, 'v')
    assert_abs(npy(pred['j']), g['pred_j'], TOL_VERT_ABS, 'j')
    assert_abs(npy(pred['poses']), g['pred_poses'], 5e-6, 'poses')
    assert_abs(npy(pred['orient']), g['pred_orient'], 2e-6, 'orient')
    assert_abs(npy(pred['trans']), g['pred_trans'], 2e-6, 'trans')
    a = m.args
    for s in range(2):
        ld, info = m.step(vi, fi)
        keys = ['gmm_loss', 'vp_recon_loss', 'vp_kl_loss', 'total_loss'] + (['instance_loss', 'loss_3d'] if version >= 3 else [])
        for k in keys:
            assert_rel(ld[k], g['s%d_%s' % (s, k)], TOL_LOSS_REL, 'step %d %s' % (s, k))
        # the fixture's kp_loss carries the CPU aliasing quirk of the reference (see test_oracle_golden.py)
        aliased = float(ld['kp_loss']) + a.weight_vp_loss * float(ld['vp_recon_loss']) + a.weight_vp_z_loss * float(ld['vp_kl_loss'])
        if version >= 3:
            aliased += a.weight_instance_loss * float(ld['instance_loss']) + a.weight_3d_loss * float(ld['loss_3d'])
        assert_rel(aliased, g['s%d_kp_loss' % s], TOL_LOSS_REL, 'step %d kp_loss' % s)
        if s == 0:
            assert_abs(npy(info['loss_all']), g['s0_loss_all'], 1e-5 * np.abs(g['s0_loss_all']).max())
            grads = m.engine.named_grads()
            for k, ref in g.items():
                if k.startswith('s0_grad.'):
                    assert_grad(npy(grads[k[8:]]), ref, what=k)
        params = m.engine.named_params()
        for k, ref in g.items():
            pre = 's%d_param.' % s
            if k.startswith(pre):
                assert_adam_param(npy(params[k[len(pre):]]), ref, k)
    # state_dict carries the reference's key names
    sd = m.state_dict()
    for k in g:
        if k.startswith('s0_param.'):
            assert k[9:] in sd, k


def test_step_vs_oracle_larger_batch():
    """N = 96 random samples (views with uneven counts), h = 64, against the CPU oracle."""
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence
    args = assets.default_args(h_dim=64, phase_rbf_dim=16, instance_code_size=4, monotonic_network_n_nodes=20,
                               batch_size=96, weight_instance_loss=0.1, weight_3d_loss=10.0, lr_human=1e-3)
    V, Fr = 5, 16
    params = assets.make_fit_params(args, V, seed=77)
    data = assets.make_fit_data(V, Fr, seed=5)
    r = np.random.RandomState(9)
    data['points2d_gt_all'][..., 0] = 540 + 200 * r.standard_normal((V, Fr, 25))
    data['points2d_gt_all'][..., 1] = 960 + 200 * r.standard_normal((V, Fr, 25))
    vi = torch.tensor(r.randint(0, V, 96)); fi = torch.tensor(r.randint(0, Fr, 96))
    orc = O.FitOracle(args, O.to_torch(params), O.to_torch(assets.make_smpl_assets(0)),
                      O.to_torch(assets.make_vposer_weights(2)), O.to_torch(assets.make_gmm(3)), O.to_torch(data),
                      version=3).build_optimizers()
    ref_ld, _ = orc.step(vi, fi)
    m = nmm.NemoV3(args, SyntheticMultiViewSequence(data=data), DEV, init_params=params)
    m.engine.compute_betas_grad = True
    ld, _ = m.step(vi, fi)
    for k in ['kp_loss', 'gmm_loss', 'vp_recon_loss', 'vp_kl_loss', 'instance_loss', 'loss_3d', 'total_loss']:
        assert_rel(ld[k], ref_ld[k], TOL_LOSS_REL, k)
    grads = m.engine.named_grads()
    for k, p in orc.p.items():
        if p.grad is not None:
            assert_grad(npy(grads[k]), npy(p.grad), what=k)
