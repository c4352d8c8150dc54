"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden            # writes tests/golden/*.npz + REPORT.txt

Inputs are seeded (numpy RandomState) and stored next to the reference outputs, so the
tests need neither the reference nor this script.  After writing, the oracle restatement is
run on the same inputs and the max deviations are printed and stored in REPORT.txt.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nemo_cvpr2023_b200 import assets  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
VSUB = np.concatenate([np.arange(0, 6890, 97), np.asarray(assets.VERTEX_JOINT_IDS)])


def T(a, grad=False):
    t = torch.tensor(np.asarray(a, dtype=np.float32))
    return t.requires_grad_(True) if grad else t


def npy(t):
    # .copy(): on CPU .numpy() aliases the tensor, and parameters / never-zeroed grads are
    # updated in place by the next step
    return t.detach().cpu().numpy().copy()


def rot_inputs(r):
    """axis-angle vectors that exercise all four quaternion branches, tiny and near-pi angles."""
    aa = r.standard_normal((40, 3))
    aa /= np.linalg.norm(aa, axis=1, keepdims=True)
    ang = np.concatenate([r.uniform(0.05, 3.1, 28), [1e-3, 1e-2, 3.0, 3.1, 3.13, 0.5, 1.0, 1.5, 2.0, 2.5, 2.9, 0.2]])
    return (aa * ang[:, None]).astype(np.float32)


def gen_geometry(mods):
    r = np.random.RandomState(100)
    out = {}
    x6 = r.standard_normal((24, 6)).astype(np.float32)
    x6[:4] *= 10.0
    x6[4:8] *= 0.01
    w9 = r.standard_normal((24, 3, 3)).astype(np.float32)
    t = T(x6, True)
    R = mods.hg.rot6d_to_rotmat(t)
    (R * T(w9)).sum().backward()
    out.update(rot6d_in=x6, rot6d_w=w9, rot6d_out=npy(R), rot6d_grad=npy(t.grad))

    aa = rot_inputs(r)
    w = r.standard_normal((40, 3, 3)).astype(np.float32)
    t = T(aa, True)
    Rq = mods.hg.batch_rodrigues(t)
    (Rq * T(w)).sum().backward()
    out.update(aa_in=aa, rod_w=w, rodq_out=npy(Rq), rodq_grad=npy(t.grad))
    t = T(aa, True)
    Rs = mods.lbs.batch_rodrigues(t)
    (Rs * T(w)).sum().backward()
    out.update(rods_out=npy(Rs), rods_grad=npy(t.grad))

    Rin = npy(Rq).copy()
    w3 = r.standard_normal((40, 3)).astype(np.float32)
    t = T(Rin, True)
    a1 = mods.hg.rotation_matrix_to_angle_axis(t)
    (a1 * T(w3)).sum().backward()
    out.update(r2aa_in=Rin, r2aa_w=w3, r2aa_hmr=npy(a1), r2aa_hmr_grad=npy(t.grad))
    t = T(Rin, True)
    a2 = mods.rt.matrot2aa(t)
    (a2 * T(w3)).sum().backward()
    out.update(r2aa_tgm=npy(a2), r2aa_tgm_grad=npy(t.grad))

    pts = (r.standard_normal((5, 7, 3)) + np.asarray([0, 0, 6.0])).astype(np.float32)
    rot = npy(mods.hg.batch_rodrigues(T(0.3 * r.standard_normal((5, 3)))))
    tr = (0.2 * r.standard_normal((5, 3))).astype(np.float32)
    fl = np.full((5,), 5000.0, np.float32)
    cc = np.tile(np.asarray([[540.0, 960.0]], np.float32), (5, 1))
    wp = r.standard_normal((5, 7, 2)).astype(np.float32)
    tp, trr, ttr = T(pts, True), T(rot, True), T(tr, True)
    p1 = mods.hg.perspective_projection(tp, trr, ttr, T(fl), T(cc))
    (p1 * T(wp)).sum().backward()
    out.update(pp_pts=pts, pp_rot=rot, pp_tr=tr, pp_fl=fl, pp_cc=cc, pp_w=wp, pp_hmr=npy(p1),
               pp_hmr_gpts=npy(tp.grad), pp_hmr_grot=npy(trr.grad), pp_hmr_gtr=npy(ttr.grad))
    p2 = mods.ug.perspective_projection(T(pts), T(rot), T(tr), T(fl), T(cc))
    K = np.zeros((5, 3, 3), np.float32)
    K[:, 0, 0] = K[:, 1, 1] = 1234.0
    K[:, 2, 2] = 1
    K[:, :2, 2] = cc
    p3 = mods.ug.perspective_projection_with_K(T(pts), T(rot), T(tr), T(K))
    out.update(pp_utils=npy(p2), pp_K=K, pp_withK=npy(p3))
    np.savez_compressed(os.path.join(GOLD, 'geometry.npz'), **out)
    return out


def gen_lbs(mods, sm):
    r = np.random.RandomState(101)
    out = {}
    n = 3
    betas = r.standard_normal((n, 10)).astype(np.float32)
    pose = (0.4 * r.standard_normal((n, 72))).astype(np.float32)
    transl = r.standard_normal((n, 3)).astype(np.float32)
    dV = r.standard_normal((n, 6890, 3)).astype(np.float32)
    dJ24 = r.standard_normal((n, 24, 3)).astype(np.float32)
    dJ49 = r.standard_normal((n, 49, 3)).astype(np.float32)
    smt = {k: torch.tensor(v) for k, v in sm.items()}
    # (a) lbs, axis-angle input, per-frame betas  (BodyModel.forward path)
    tb, tp = T(betas, True), T(pose, True)
    v, j = mods.lbs.lbs(tb, tp, smt['v_template'], smt['shapedirs'], smt['posedirs'], smt['J_regressor'],
                        smt['parents'], smt['lbs_weights'], pose2rot=True)
    ((v * T(dV)).sum() + (j * T(dJ24)).sum()).backward()
    out.update(betas=betas, pose=pose, transl=transl, dV=dV, dJ24=dJ24, dJ49=dJ49,
               lbs_v_sub=npy(v)[:, VSUB], lbs_v_sum=npy(v).astype(np.float64).sum((1, 2)),
               lbs_v_abs=np.abs(npy(v).astype(np.float64)).sum((1, 2)), lbs_j=npy(j),
               lbs_gbetas=npy(tb.grad), lbs_gpose=npy(tp.grad))
    # (b) hmr.smpl.SMPL.forward with rotation matrices, shared betas (fit path)
    smpl = mods.hs.SMPL('unused', batch_size=1, create_transl=False)
    rot = npy(mods.hg.batch_rodrigues(T(pose.reshape(-1, 3)))).reshape(n, 24, 3, 3)
    b1 = betas[:1].copy()
    tb, tr = T(b1, True), T(rot, True)
    o = smpl(betas=tb, body_pose=tr[:, 1:], global_orient=tr[:, :1], pose2rot=False)
    ((o.vertices * T(dV)).sum() + (o.joints * T(dJ49)).sum()).backward()
    out.update(rotmats=rot, betas1=b1, smpl_v_sub=npy(o.vertices)[:, VSUB],
               smpl_v_sum=npy(o.vertices).astype(np.float64).sum((1, 2)), smpl_j49=npy(o.joints),
               smpl_gbetas=npy(tb.grad), smpl_grot=npy(tr.grad))
    # (c) SMPL.forward with axis-angle + transl (eval path: pose2rot=True)
    tp = T(pose, True)
    o = smpl(betas=T(betas), body_pose=tp[:, 3:], global_orient=tp[:, :3], transl=T(transl), pose2rot=True)
    ((o.vertices * T(dV)).sum() + (o.joints * T(dJ49)).sum()).backward()
    out.update(smplaa_v_sub=npy(o.vertices)[:, VSUB], smplaa_j49=npy(o.joints), smplaa_gpose=npy(tp.grad))
    # (d) defaults: SMPL.forward() with everything None -> rest pose
    o = smpl()
    out.update(rest_v_sub=npy(o.vertices)[:, VSUB], rest_j49=npy(o.joints))
    np.savez_compressed(os.path.join(GOLD, 'lbs.npz'), **out)
    return out


def gen_modules(mods, vp_np, gmm_np):
    r = np.random.RandomState(102)
    out = {}
    x = r.uniform(0, 1, (9, 1)).astype(np.float32)
    ls = (0.3 * r.standard_normal(12)).astype(np.float32)
    names = ['gaussian', 'linear', 'quadratic', 'inverse_quadratic', 'multiquadric',
             'inverse_multiquadric', 'spline', 'poisson_one', 'poisson_two', 'matern32', 'matern52']
    out.update(rbf_x=x, rbf_ls=ls)
    for nm in names:
        layer = mods.rbf.RBF(12, nm)
        layer.log_sigmas.data = T(ls)
        tx = T(x, True)
        y = layer(tx)
        y.sum().backward()
        out['rbf_' + nm] = npy(y)
        out['rbf_' + nm + '_gx'] = npy(tx.grad)
        out['rbf_' + nm + '_gls'] = npy(layer.log_sigmas.grad)
    mono = mods.mono.MonotonicNetwork(7, 'linear')
    sh = r.uniform(-0.1, 1, 7).astype(np.float32)
    sc = r.uniform(-2, 20, 7).astype(np.float32)
    mono.shifts.data, mono.scales.data = T(sh), T(sc)
    tx = T(x, True)
    wy = r.standard_normal((9, 1)).astype(np.float32)
    y = mono(tx)
    (y * T(wy)).sum().backward()
    out.update(mono_sh=sh, mono_sc=sc, mono_w=wy, mono_y=npy(y), mono_gx=npy(tx.grad),
               mono_gsh=npy(mono.shifts.grad), mono_gsc=npy(mono.scales.grad))
    # GMoF + keypoint losses
    pred = r.uniform(0, 1500, (6, 25, 2)).astype(np.float32)
    gt = r.uniform(0, 1500, (6, 25, 2)).astype(np.float32)
    conf = r.uniform(0, 1, (6, 25, 1)).astype(np.float32)
    size = r.uniform(200, 700, (6,)).astype(np.float32)
    host = mods.nmm.MultiViewModel.__new__(mods.nmm.MultiViewModel)
    torch.nn.Module.__init__(host)
    host.criterion_keypoints = torch.nn.MSELoss(reduction='none')
    host.robustifier = mods.mu.GMoF()
    out.update(kl_pred=pred, kl_gt=gt, kl_conf=conf, kl_size=size)
    for lt in ['rmse', 'rmse_resized', 'mse', 'rmse_robust', 'mse_robust', 'mse_robust_resized']:
        tp = T(pred, True)
        l = host.keypoint_loss(tp, T(gt), T(conf), T(size), loss_type=lt)
        l.sum().backward()
        out['kl_' + lt] = npy(l)
        out['kl_' + lt + '_g'] = npy(tp.grad)
    # VPoser
    vp = rh.build_vposer(mods, vp_np)
    pb = (0.3 * r.standard_normal((5, 63))).astype(np.float32)
    tp = T(pb, True)
    q = vp.encode(tp)
    dec = vp.decode(q.mean)
    wd = r.standard_normal((5, 21, 3)).astype(np.float32)
    kl = torch.mean(torch.sum(torch.distributions.kl.kl_divergence(
        q, torch.distributions.normal.Normal(torch.zeros(5, 32), torch.ones(5, 32))), dim=[1]))
    ((dec['pose_body'] * T(wd)).sum() + kl).backward()
    out.update(vp_in=pb, vp_mu=npy(q.mean), vp_sigma=npy(q.scale), vp_dec=npy(dec['pose_body']),
               vp_matrot=npy(dec['pose_body_matrot']), vp_w=wd, vp_kl=npy(kl), vp_gin=npy(tp.grad))
    # GMM prior
    pr = rh.build_prior(mods, gmm_np)
    ps = (0.3 * r.standard_normal((6, 69))).astype(np.float32)
    tp = T(ps, True)
    ll = pr(tp, None)
    ll.sum().backward()
    out.update(gmm_in=ps, gmm_ll=npy(ll), gmm_g=npy(tp.grad))
    np.savez_compressed(os.path.join(GOLD, 'modules.npz'), **out)
    return out


STEP_CFG = dict(h_dim=48, phase_rbf_dim=12, instance_code_size=3, monotonic_network_n_nodes=7,
                batch_size=10, lr_human=1e-3)


def step_case(version):
    over = dict(STEP_CFG)
    if version >= 3:
        over.update(weight_instance_loss=0.5, weight_3d_loss=1000.0)
    args = assets.default_args(**over)
    V, Fr = 3, 6
    params = assets.make_fit_params(args, V, seed=40 + version)
    data = assets.make_fit_data(V, Fr, seed=1)
    r = np.random.RandomState(200 + version)
    # keypoints near the actual projections so the robust loss is in its curved region
    data['points2d_gt_all'][..., 0] = 540 + 300 * r.standard_normal((V, Fr, 25))
    data['points2d_gt_all'][..., 1] = 960 + 300 * r.standard_normal((V, Fr, 25))
    data['hmr_mask'][0, 1] = 0.0
    view_idx = np.asarray([0, 0, 2, 2, 2, 0, 2, 0, 2, 2], np.int64)   # view 1 absent on purpose
    frame_idx = np.asarray([0, 1, 5, 3, 3, 4, 0, 2, 1, 2], np.int64)
    return args, params, data, view_idx, frame_idx


def gen_step(mods, sm, vp_np, gmm_np, version):
    args, params, data, view_idx, frame_idx = step_case(version)
    m = rh.build_fit_model(mods, args, params, sm, vp_np, gmm_np, data, version=version)
    out = dict(view_idx=view_idx, frame_idx=frame_idx)
    vi, fi = torch.tensor(view_idx), torch.tensor(frame_idx)
    pred = m.get_preds_batch(vi, fi)
    out.update(pred_v_sub=npy(pred['v'])[:, VSUB], pred_j=npy(pred['j']), pred_poses=npy(pred['poses']),
               pred_orient=npy(pred['orient']), pred_orient_aa=npy(pred['orient_aa']),
               pred_trans=npy(pred['trans']))
    for s in range(2):
        loss_dict, info = m.step(vi, fi)
        for k, v in loss_dict.items():
            out['s%d_%s' % (s, k)] = np.asarray(v, dtype=np.float64)
        if s == 0:
            out['s0_loss_all'] = npy(info['loss_all'])
            for k, p in m.named_parameters():
                if p.grad is not None:
                    out['s0_grad.' + k] = npy(p.grad)
        for k, p in m.named_parameters():
            if p.requires_grad:
                out['s%d_param.%s' % (s, k)] = npy(p)
    np.savez_compressed(os.path.join(GOLD, 'step_v%d.npz' % version), **out)
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(1)   # fixed summation order for the fixtures
    sm = assets.make_smpl_assets(0)
    vp_np = assets.make_vposer_weights(2)
    gmm_np = assets.make_gmm(3)
    mods = rh.setup(sm)
    gen_geometry(mods)
    gen_lbs(mods, sm)
    gen_modules(mods, vp_np, gmm_np)
    for v in (2, 3, 4):
        gen_step(mods, sm, vp_np, gmm_np, v)
    os.chdir(ROOT)
    print('golden fixtures written to', GOLD)
    for f in sorted(os.listdir(GOLD)):
        print('  %-16s %8.1f KB' % (f, os.path.getsize(os.path.join(GOLD, f)) / 1024))


if __name__ == '__main__':
    main()
