"""Pin the oracle restatement against outputs of the unmodified reference
(tests/golden/*.npz, produced by oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from nemo_cvpr2023_b200 import assets
from oracle import nemo_oracle as O
from oracle.make_golden import VSUB, step_case
from tests.util import (T, TOL_GRAD_REL, TOL_LOSS_REL, TOL_VERT_ABS, assert_abs, assert_grad,
                        assert_rel, gold, npy)

torch.set_num_threads(4)


@pytest.fixture(scope='module')
def sm():
    return O.to_torch(assets.make_smpl_assets(0))


@pytest.fixture(scope='module')
def vp():
    return O.to_torch(assets.make_vposer_weights(2))


@pytest.fixture(scope='module')
def gmm():
    return O.to_torch(assets.make_gmm(3))


def test_rot6d():
    g = gold('geometry')
    x = T(g['rot6d_in'], True)
    r = O.rot6d_to_rotmat(x)
    (r * T(g['rot6d_w'])).sum().backward()
    assert_abs(npy(r), g['rot6d_out'], 1e-6, 'rot6d')
    assert_grad(npy(x.grad), g['rot6d_grad'], what='rot6d grad')


@pytest.mark.parametrize('kind', ['quat', 'skew'])
def test_rodrigues(kind):
    g = gold('geometry')
    x = T(g['aa_in'], True)
    fn = O.batch_rodrigues_quat if kind == 'quat' else O.batch_rodrigues_skew
    key = 'rodq' if kind == 'quat' else 'rods'
    r = fn(x)
    (r * T(g['rod_w'])).sum().backward()
    assert_abs(npy(r), g[key + '_out'], 1e-6, kind)
    assert_grad(npy(x.grad), g[key + '_grad'], what=kind + ' grad')


@pytest.mark.parametrize('kind', ['hmr', 'tgm'])
def test_rotmat_to_aa(kind):
    g = gold('geometry')
    x = T(g['r2aa_in'], True)
    aa = O.rotation_matrix_to_angle_axis(x, nan_to_zero=(kind == 'hmr'))
    (aa * T(g['r2aa_w'])).sum().backward()
    assert_abs(npy(aa), g['r2aa_' + kind], 2e-6, kind)
    assert_grad(npy(x.grad), g['r2aa_%s_grad' % kind], what=kind + ' grad')
    # round trip with the quaternion Rodrigues for angles < pi
    assert_abs(npy(aa), g['aa_in'], 2e-4, 'aa round trip')


def test_projection():
    g = gold('geometry')
    p, r, t = T(g['pp_pts'], True), T(g['pp_rot'], True), T(g['pp_tr'], True)
    o = O.perspective_projection(p, r, t, T(g['pp_fl']), T(g['pp_cc']))
    (o * T(g['pp_w'])).sum().backward()
    assert_abs(npy(o), g['pp_hmr'], 2e-3, 'proj')       # pixels, magnitude ~1e3
    assert_grad(npy(p.grad), g['pp_hmr_gpts'])
    assert_grad(npy(r.grad), g['pp_hmr_grot'])
    assert_grad(npy(t.grad), g['pp_hmr_gtr'])
    o2 = O.perspective_projection(T(g['pp_pts']), T(g['pp_rot']), T(g['pp_tr']), T(g['pp_fl']),
                                  T(g['pp_cc']), epsilon=1e-8)
    assert_abs(npy(o2), g['pp_utils'], 2e-3)
    o3 = O.perspective_projection_with_K(T(g['pp_pts']), T(g['pp_rot']), T(g['pp_tr']), T(g['pp_K']))
    assert_abs(npy(o3), g['pp_withK'], 2e-3)
    # analytic identity: a point on the optical axis projects to the principal point
    pt = torch.tensor([[[0.0, 0.0, 3.0]]])
    o4 = O.perspective_projection(pt, torch.eye(3)[None], torch.zeros(1, 3), 5000.0,
                                  torch.tensor([[540.0, 960.0]]))
    assert_abs(npy(o4), [[[540.0, 960.0]]], 0)


def test_lbs_axis_angle(sm):
    g = gold('lbs')
    b, p = T(g['betas'], True), T(g['pose'], True)
    v, j = O.lbs(b, p, sm['v_template'], sm['shapedirs'], sm['posedirs'], sm['J_regressor'],
                 sm['parents'], sm['lbs_weights'], pose2rot=True)
    ((v * T(g['dV'])).sum() + (j * T(g['dJ24'])).sum()).backward()
    assert_abs(npy(v)[:, VSUB], g['lbs_v_sub'], TOL_VERT_ABS, 'lbs verts')
    assert_abs(npy(v).sum((1, 2)), g['lbs_v_sum'], 2e-3, 'lbs vert checksum')
    assert_abs(npy(j), g['lbs_j'], TOL_VERT_ABS, 'lbs joints')
    assert_grad(npy(b.grad), g['lbs_gbetas'], what='d betas')
    assert_grad(npy(p.grad), g['lbs_gpose'], what='d pose')
    # BodyModel.forward = lbs + trans
    bm = O.body_model_forward(sm, root_orient=T(g['pose'][:, :3]), pose_body=T(g['pose'][:, 3:66]),
                              pose_hand=T(g['pose'][:, 66:]), betas=T(g['betas']), trans=T(g['transl']))
    assert_abs(npy(bm['v'])[:, VSUB], g['lbs_v_sub'] + g['transl'][:, None], TOL_VERT_ABS)
    assert_abs(npy(bm['Jtr']), g['lbs_j'] + g['transl'][:, None], TOL_VERT_ABS)


def test_smpl_forward(sm):
    g = gold('lbs')
    b, r = T(g['betas1'], True), T(g['rotmats'], True)
    o = O.smpl_forward(sm, betas=b, body_pose=r[:, 1:], global_orient=r[:, :1], pose2rot=False)
    ((o.vertices * T(g['dV'])).sum() + (o.joints * T(g['dJ49'])).sum()).backward()
    assert_abs(npy(o.vertices)[:, VSUB], g['smpl_v_sub'], TOL_VERT_ABS)
    assert_abs(npy(o.joints), g['smpl_j49'], TOL_VERT_ABS)
    assert_grad(npy(b.grad), g['smpl_gbetas'])
    assert_grad(npy(r.grad), g['smpl_grot'])
    p = T(g['pose'], True)
    o = O.smpl_forward(sm, betas=T(g['betas']), body_pose=p[:, 3:], global_orient=p[:, :3],
                       transl=T(g['transl']), pose2rot=True)
    ((o.vertices * T(g['dV'])).sum() + (o.joints * T(g['dJ49'])).sum()).backward()
    assert_abs(npy(o.vertices)[:, VSUB], g['smplaa_v_sub'], TOL_VERT_ABS)
    assert_abs(npy(o.joints), g['smplaa_j49'], TOL_VERT_ABS)
    assert_grad(npy(p.grad), g['smplaa_gpose'])
    o = O.smpl_forward(sm)
    assert_abs(npy(o.vertices)[:, VSUB], g['rest_v_sub'], TOL_VERT_ABS)
    assert_abs(npy(o.joints), g['rest_j49'], TOL_VERT_ABS)
    # analytic identity: rest pose, zero betas => vertices == v_template
    assert_abs(npy(o.vertices)[0], npy(sm['v_template']), 2e-6)


def test_rbf_all_bases():
    g = gold('modules')
    names = ['gaussian', 'linear', 'quadratic', 'inverse_quadratic', 'multiquadric',
             'inverse_multiquadric', 'spline', 'poisson_one', 'poisson_two', 'matern32', 'matern52']
    c = torch.linspace(0, 1, 12)[:, None]
    for nm in names:
        x, ls = T(g['rbf_x'], True), T(g['rbf_ls'], True)
        y = O.rbf_forward(x, c, ls, nm)
        y.sum().backward()
        assert_abs(npy(y), g['rbf_' + nm], 2e-6, nm)
        assert_grad(npy(x.grad), g['rbf_%s_gx' % nm], what=nm)
        assert_grad(npy(ls.grad), g['rbf_%s_gls' % nm], what=nm)


def test_monotonic():
    g = gold('modules')
    x, sh, sc = T(g['rbf_x'], True), T(g['mono_sh'], True), T(g['mono_sc'], True)
    y = O.monotonic_forward(x, sh, sc)
    (y * T(g['mono_w'])).sum().backward()
    assert_abs(npy(y), g['mono_y'], 2e-6)
    assert_grad(npy(x.grad), g['mono_gx'])
    assert_grad(npy(sh.grad), g['mono_gsh'])
    assert_grad(npy(sc.grad), g['mono_gsc'])


def test_keypoint_losses():
    g = gold('modules')
    for lt in ['rmse', 'rmse_resized', 'mse', 'rmse_robust', 'mse_robust', 'mse_robust_resized']:
        p = T(g['kl_pred'], True)
        l = O.keypoint_loss(p, T(g['kl_gt']), T(g['kl_conf']), T(g['kl_size']), lt)
        l.sum().backward()
        scale = max(1.0, np.abs(g['kl_' + lt]).max())
        assert_abs(npy(l), g['kl_' + lt], 1e-5 * scale, lt)
        assert_grad(npy(p.grad), g['kl_%s_g' % lt], what=lt)


def test_vposer(vp):
    g = gold('modules')
    x = T(g['vp_in'], True)
    mu, sigma = O.vposer_encode(vp, x)
    dec, matrot = O.vposer_decode(vp, mu)
    kl = (-torch.log(sigma) + 0.5 * (sigma ** 2 + mu ** 2) - 0.5).sum(1).mean()
    ((dec * T(g['vp_w'])).sum() + kl).backward()
    assert_abs(npy(mu), g['vp_mu'], 2e-6)
    assert_abs(npy(sigma), g['vp_sigma'], 2e-6)
    assert_abs(npy(dec), g['vp_dec'], 5e-6)
    assert_abs(npy(matrot), g['vp_matrot'], 2e-6)
    assert_rel(kl, g['vp_kl'], TOL_LOSS_REL)
    assert_grad(npy(x.grad), g['vp_gin'])


def test_gmm(gmm):
    g = gold('modules')
    x = T(g['gmm_in'], True)
    ll = O.gmm_prior(gmm, x)
    ll.sum().backward()
    assert_abs(npy(ll) / g['gmm_ll'], np.ones_like(g['gmm_ll']), TOL_LOSS_REL)
    assert_grad(npy(x.grad), g['gmm_g'])


def build_oracle_model(version, sm, vp, gmm, dtype=torch.float32):
    args, params, data, view_idx, frame_idx = step_case(version)
    p = O.to_torch(params, dtype)
    d = O.to_torch(data, dtype)
    m = O.FitOracle(args, p, sm, vp, gmm, d, version=version)
    return m.build_optimizers(), torch.tensor(view_idx), torch.tensor(frame_idx)


@pytest.mark.parametrize('version', [2, 3, 4])
def test_step_matches_reference(version, sm, vp, gmm):
    g = gold('step_v%d' % version)
    m, vi, fi = build_oracle_model(version, sm, vp, gmm)
    with torch.no_grad():
        pred = m.get_preds_batch(vi, fi)
    assert_abs(npy(pred['v'])[:, VSUB], g['pred_v_sub'], TOL_VERT_ABS, 'v')
    assert_abs(npy(pred['j']), g['pred_j'], TOL_VERT_ABS, 'j')
    assert_abs(npy(pred['poses']), g['pred_poses'], 5e-6, 'poses')
    assert_abs(npy(pred['trans']), g['pred_trans'], 2e-6, 'trans')
    for s in range(2):
        ld, info = m.step(vi, fi)
        for k in ['gmm_loss', 'vp_recon_loss', 'vp_kl_loss', 'total_loss'] + (
                ['instance_loss', 'loss_3d'] if version >= 3 else []):
            assert_rel(ld[k], g['s%d_%s' % (s, k)], TOL_LOSS_REL, 'step %d %s' % (s, k))
        # Reference quirk, CPU only: loss_dict['kp_loss'] = loss.detach().cpu().numpy() ALIASES `loss`
        # on a CPU run (nemo/neural_motion_model.py:3564), and the following `loss += ...` lines
        # (:3571-3574, :3861-3882) are in place, so the fixture's kp_loss also contains those terms
        # (everything before the out-of-place `loss = loss + gmm`, :3579).  On CUDA .cpu() copies.
        a = m.args
        aliased = ld['kp_loss'] + a.weight_vp_loss * ld['vp_recon_loss'] + a.weight_vp_z_loss * ld['vp_kl_loss']
        if version >= 3:
            aliased += a.weight_instance_loss * ld['instance_loss'] + a.weight_3d_loss * ld['loss_3d']
        assert_rel(aliased, g['s%d_kp_loss' % s], TOL_LOSS_REL, 'step %d kp_loss' % s)
        if s == 0:
            assert_abs(npy(info['loss_all']), g['s0_loss_all'], 1e-5 * np.abs(g['s0_loss_all']).max())
            for k, ref in g.items():
                if k.startswith('s0_grad.'):
                    assert_grad(npy(m.p[k[8:]].grad), ref, what=k)
        for k, ref in g.items():
            pre = 's%d_param.' % s
            if k.startswith(pre):
                # Adam's first steps move each weight by ~lr regardless of gradient scale
                assert_abs(npy(m.p[k[len(pre):]]), ref, 2e-5 * max(1.0, np.abs(ref).max()), k)


def test_metrics_oracle():
    """oracle/pose_metrics.py == the reference's pose_utils / f_pck outputs (tests/golden/metrics.npz)."""
    from oracle import pose_metrics as M
    g = gold('metrics')
    for name in ('j', 'v', 'refl'):
        S1, S2 = g[name + '_S1'].astype(np.float64), g[name + '_S2'].astype(np.float64)
        hat = np.stack([M.compute_similarity_transform(a, b)[0] for a, b in zip(S1, S2)])
        assert np.abs(hat - g[name + '_hat']).max() < 1e-6
        if name != 'refl':
            assert np.allclose(M.reconstruction_error(S1, S2, pa=True, reduction=None), g[name + '_err_pa'], rtol=1e-9)
            assert np.allclose(M.reconstruction_error(S1, S2, pa=False, reduction=None), g[name + '_err'], rtol=1e-12)
            _, s, R, t = M.compute_similarity_transform(S1[0], S2[0])
            assert np.allclose(np.concatenate([[s], R.reshape(-1), t]), g[name + '_srt0'], atol=1e-9)
    assert abs(M.pck(g['pck_pred'].astype(np.float64), g['pck_gt'].astype(np.float64), g['pck_size'].astype(np.float64))
               - float(g['pck'])) < 1e-4
