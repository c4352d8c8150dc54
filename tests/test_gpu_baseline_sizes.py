"""Parity of the BENCHMARKED code path at BASELINE.json sizes.

At these sizes every large GEMM of the step dispatches to the tcgen05 3xTF32 kernel (forward, dgrad with the fused
activation gate, split-K wgrad at K = 513 / 4097, the K = 20 736 pose-blend dgrad), which the small golden cases never
reach -- there M.N.K < 2e6 sends them to the CUDA-core kernel.  Checked against the CPU oracle on identical inputs
with the contract tolerances (losses 1e-5 relative, gradients 1e-4 relative per tensor, vertices / joints 1e-5 m).
Also here: the five keypoint-loss modes the round-1 tests did not reach through the fused kernel, the plateau
scheduler (`lr_factor < 1`), a checkpoint WRITTEN BY THE REFERENCE, and the warm-up -> fit hand-over of Adam state.
"""
import os

import numpy as np
import pytest
import torch

from nemo_cvpr2023_b200 import assets
from oracle import nemo_oracle as O
from oracle.make_golden import step_case
from tests.util import (GOLD, TOL_GRAD_REL, TOL_LOSS_REL, TOL_VERT_ABS, assert_abs, assert_adam_param, assert_grad,
                        assert_rel, npy)

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _oracle(args, params, data, version):
    return O.FitOracle(args, O.to_torch(params), O.to_torch(assets.make_smpl_assets(0)),
                       O.to_torch(assets.make_vposer_weights(2)), O.to_torch(assets.make_gmm(3)), O.to_torch(data),
                       version=version).build_optimizers()


def _model(args, data, params, version):
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence
    cls = {2: nmm.NemoV2, 3: nmm.NemoV3, 4: nmm.NemoV4}[version]
    return cls(args, SyntheticMultiViewSequence(data=data), DEV, init_params=params)


def _near(data, V, Fr, seed, sd):
    """keypoints around the image centre, so the robust loss works in its curved region"""
    r = np.random.RandomState(seed)
    data['points2d_gt_all'][..., 0] = 540 + sd * r.standard_normal((V, Fr, 25))
    data['points2d_gt_all'][..., 1] = 960 + sd * r.standard_normal((V, Fr, 25))


def _compare_step(orc, m, vi, fi, keys, skip=('learned_betas',)):
    ref_ld, ref_info = orc.step(vi, fi)
    ld, info = m.step(vi, fi)
    for k in keys:
        assert_rel(ld[k], ref_ld[k], TOL_LOSS_REL, k)
    grads = m.engine.named_grads()
    n = 0
    for k, p in orc.p.items():
        if p.grad is None or k in skip:
            continue
        assert_grad(npy(grads[k]), npy(p.grad), what=k)
        n += 1
    return ld, info, ref_info, n


@pytest.mark.parametrize('version', [2, 3])
def test_configs0_step_matches_oracle(version):
    """BASELINE configs[0]: 1 action x 8 video instances x 64 frames = 512 samples, motion MLP h = 1000, RBF 100,
    instance code 5, VPoser v2v + KL, GMM prior, mse_robust 2D loss (+ 3D pose loss and instance regulariser for V3).
    V2 runs on the bench's own data (uniform keypoints); V3 with keypoints near the projections."""
    V, Fr = 8, 64
    over = dict(batch_size=V * Fr)
    if version >= 3:
        over.update(weight_instance_loss=0.1, weight_3d_loss=1000.0)
    args = assets.default_args(**over)
    params = assets.make_fit_params(args, V, seed=4)
    data = assets.make_fit_data(V, Fr, seed=1)
    if version >= 3:
        _near(data, V, Fr, 7, 250.0)
    vi, fi = torch.arange(V).repeat_interleave(Fr), torch.arange(Fr).repeat(V)
    orc, m = _oracle(args, params, data, version), _model(args, data, params, version)
    keys = ['kp_loss', 'gmm_loss', 'vp_recon_loss', 'vp_kl_loss', 'total_loss']
    if version >= 3:
        keys += ['instance_loss', 'loss_3d']
    ld, info, ref_info, n = _compare_step(orc, m, vi, fi, keys)
    assert n >= 29          # cameras, instance codes, 11 motion / RBF tensors, 16 phase tensors
    ref_all = npy(ref_info['loss_all'])
    assert_abs(npy(info['loss_all']), ref_all, 1e-5 * np.abs(ref_all).max(), 'loss_all')
    # parameters after Adam (step 1) and a second step on the updated parameters
    for k, p in orc.p.items():
        if p.grad is not None and k != 'learned_betas':
            assert_adam_param(npy(m.engine.named_params()[k]), npy(p), k)
    ref2, _ = orc.step(vi, fi)
    ld2, _ = m.step(vi, fi)
    # one Adam step moves every weight by ~lr whatever its gradient: fp32 noise in near-zero gradients shows up in
    # the second step's loss at the 1e-4 level, so this is a sanity bound, not the parity contract
    assert_rel(ld2['total_loss'], ref2['total_loss'], 2e-3, 'total_loss after one update')


@pytest.mark.parametrize('per_frame_betas', [False, True])
def test_smpl_forward_backward_4096(per_frame_betas):
    """BASELINE configs[1]: SMPL forward + backward, batch 4096 poses, `betas` (1,10) and (4096,10), upstream
    dV ~ N(0,1), dJ ~ N(0,1): vertices / joints of a 64-frame subsample, d rotmats of the subsample and d betas
    (per frame: the subsample; shared: the sum over all 4096 frames, oracle evaluated in chunks)."""
    from nemo_cvpr2023_b200.hmr.smpl import SMPL
    N, SUB = 4096, 64
    g = torch.Generator().manual_seed(5)
    aa = 0.5 * torch.randn(N * 24, 3, generator=g)
    rot = O.batch_rodrigues_quat(aa).view(N, 24, 3, 3)
    betas = torch.randn(N if per_frame_betas else 1, 10, generator=g)
    dJ = torch.randn(N, 49, 3, generator=g)
    dV = torch.randn(N, 6890, 3, generator=g)
    smpl = SMPL('synthetic')
    b, r = betas.to(DEV).requires_grad_(True), rot.to(DEV).requires_grad_(True)
    out = smpl(betas=b, body_pose=r[:, 1:], global_orient=r[:, :1], pose2rot=False)
    ((out.vertices * dV.to(DEV)).sum() + (out.joints * dJ.to(DEV)).sum()).backward()
    sm = O.to_torch(assets.make_smpl_assets(0))
    sub = torch.arange(0, N, N // SUB)
    chunks = [sub] if per_frame_betas else [torch.arange(i, i + 512) for i in range(0, N, 512)]
    gb_sum = torch.zeros(1, 10, dtype=torch.float64)
    torch.set_num_threads(os.cpu_count() or 1)
    for idx in chunks:
        bo = (betas[idx] if per_frame_betas else betas).clone().requires_grad_(True)
        ro = rot[idx].clone().requires_grad_(True)
        ref = O.smpl_forward(sm, betas=bo, body_pose=ro[:, 1:], global_orient=ro[:, :1], pose2rot=False)
        ((ref.vertices * dV[idx]).sum() + (ref.joints * dJ[idx]).sum()).backward()
        if per_frame_betas:
            assert_abs(npy(out.vertices[idx.to(DEV)]), npy(ref.vertices), TOL_VERT_ABS, 'vertices')
            assert_abs(npy(out.joints[idx.to(DEV)]), npy(ref.joints), TOL_VERT_ABS, 'joints')
            assert_grad(npy(r.grad[idx.to(DEV)]), npy(ro.grad), what='d rotmats')
            assert_grad(npy(b.grad[idx.to(DEV)]), npy(bo.grad), what='d betas (per frame)')
        else:
            gb_sum += bo.grad.double()
            take = torch.isin(idx, sub)
            if take.any():
                loc = torch.nonzero(take).reshape(-1)
                gl = idx[loc].to(DEV)
                assert_abs(npy(out.vertices[gl]), npy(ref.vertices[loc]), TOL_VERT_ABS, 'vertices')
                assert_abs(npy(out.joints[gl]), npy(ref.joints[loc]), TOL_VERT_ABS, 'joints')
                assert_grad(npy(r.grad[gl]), npy(ro.grad[loc]), what='d rotmats')
    if not per_frame_betas:
        assert_grad(npy(b.grad), gb_sum.numpy(), what='d betas (shared, summed over 4096 frames)')


@pytest.mark.parametrize('loss', ['rmse', 'rmse_resized', 'mse', 'rmse_robust', 'mse_robust_resized'])
def test_keypoint_loss_modes_through_the_fused_kernel(loss):
    """MultiViewModel.keypoint_loss modes (nemo/neural_motion_model.py:2806-2843) inside `step`: value, loss_all and
    every gradient through k_kp_loss_fwd / k_kp_loss_bwd."""
    V, Fr, N = 5, 16, 96
    args = assets.default_args(h_dim=64, phase_rbf_dim=16, instance_code_size=4, monotonic_network_n_nodes=20,
                               batch_size=N, loss=loss)
    params = assets.make_fit_params(args, V, seed=70)
    data = assets.make_fit_data(V, Fr, seed=6)
    _near(data, V, Fr, 8, 120.0)
    # the models derive the box size from the labels themselves (collate_gt_2d, :2908-2961): diagonal of the keypoints'
    # bounding box + 1e-4; hand the oracle the same numbers
    pts = data['points2d_gt_all']
    d0, d1 = pts[..., 0].max(-1) - pts[..., 0].min(-1), pts[..., 1].max(-1) - pts[..., 1].min(-1)
    data['gt_bbox_size'] = (np.sqrt(d0 ** 2 + d1 ** 2) + 1e-4).astype(np.float32)
    r = np.random.RandomState(10)
    vi, fi = torch.tensor(r.randint(0, V, N)), torch.tensor(r.randint(0, Fr, N))
    orc, m = _oracle(args, params, data, 2), _model(args, data, params, 2)
    ld, info, ref_info, n = _compare_step(orc, m, vi, fi, ['kp_loss', 'total_loss'])
    ref_all = npy(ref_info['loss_all'])
    got = npy(info['loss_all']).reshape(-1)[:ref_all.size].reshape(ref_all.shape)
    assert_abs(got, ref_all, 1e-5 * np.abs(ref_all).max(), 'loss_all (%s)' % loss)


def test_plateau_scheduler_matches_torch():
    """`lr_factor < 1` (:3433-3438): one ReduceLROnPlateau(factor, min_lr=1e-6) per optimiser, stepped with the total
    loss after every update (:3590-3592).  Both sides get patience 0 and an unreachable improvement threshold, so every
    step after the first halves every learning rate whatever the loss does: the fused Adam must pick each new rate up
    (losses and parameters follow torch.optim.Adam on the oracle)."""
    V, Fr, N = 3, 8, 24
    args = assets.default_args(h_dim=40, phase_rbf_dim=10, instance_code_size=3, monotonic_network_n_nodes=12,
                               batch_size=N, lr_factor=0.5, lr_human=1e-3)
    params = assets.make_fit_params(args, V, seed=51)
    data = assets.make_fit_data(V, Fr, seed=52)
    _near(data, V, Fr, 53, 150.0)
    r = np.random.RandomState(54)
    vi, fi = torch.tensor(r.randint(0, V, N)), torch.tensor(r.randint(0, Fr, N))
    orc, m = _oracle(args, params, data, 2), _model(args, data, params, 2)
    sched = [torch.optim.lr_scheduler.ReduceLROnPlateau(o, factor=args.lr_factor, min_lr=1e-6) for o in orc.optimizers]
    assert len(m.schedulers) == len(m.optimizers) == len(sched)
    for s in list(m.schedulers) + sched:
        assert s.factor == 0.5 and s.min_lrs == [1e-6]
        s.patience, s.threshold = 0, 0.9
    for it in range(5):
        ref, _ = orc.step(vi, fi)
        for s in sched:
            s.step(ref['total_loss'])
        ld, _ = m.step(vi, fi)
        assert_rel(ld['total_loss'], ref['total_loss'], 1e-5 if it == 0 else 5e-3, 'total_loss at step %d' % it)
        for o, ro in zip(m.optimizers, orc.optimizers):
            assert o.param_groups[0]['lr'] == pytest.approx(ro.param_groups[0]['lr'], rel=1e-12), it
    assert m.optimizers[1].param_groups[0]['lr'] == pytest.approx(1e-3 / 16)
    for k, p in orc.p.items():
        if p.grad is not None and k != 'learned_betas':
            assert_adam_param(npy(m.engine.named_params()[k]), npy(p), k, lr=5 * (args.lr_camera if 'camera' in k else 1e-3))


def test_loads_a_checkpoint_written_by_the_reference():
    """tests/golden/ref_ckpt_v3.pt was written by the unmodified reference's `save` after its own warm-up (3) and two fit
    steps (oracle/make_golden_ckpt.py); `load` here must put every tensor, Adam moment and per-parameter step count
    where the engine's flat buffers keep them: parameters after load equal the reference's, and so does the NEXT step
    (loss and updated parameters), which depends on exp_avg / exp_avg_sq / step of every tensor."""
    g = dict(np.load(os.path.join(GOLD, 'ref_ckpt_v3.npz')))
    args, _, data, view_idx, frame_idx = step_case(3)
    m = _model(args, data, assets.make_fit_params(args, 3, seed=1234), 3)
    m.load(os.path.join(GOLD, 'ref_ckpt_v3.pt'))
    params = m.engine.named_params()
    for k, ref in g.items():
        if k.startswith('loaded_param.'):
            assert_abs(npy(params[k[13:]]), ref, 0.0, k)
    e = m.engine
    assert int(e.g_motion.step_dev) == 5 and int(e.g_motion.step_lag_dev) == 2   # warm-up never stepped linear_out.*
    assert int(e.g_cam.step_dev) == 2 and int(e.g_phase.step_dev) == 5 and int(e.g_inst.step_dev) == 2
    ld, _ = m.step(torch.tensor(view_idx), torch.tensor(frame_idx))
    for k in ['gmm_loss', 'vp_recon_loss', 'vp_kl_loss', 'instance_loss', 'loss_3d', 'total_loss']:
        assert_rel(ld[k], g['loss.' + k], TOL_LOSS_REL, k)
    for k, ref in g.items():
        if k.startswith('after_param.'):
            name = k[12:]
            lr = args.lr_camera if name == 'learned_cameras' else 1e-3
            assert_adam_param(npy(params[name]), ref, k, lr=lr)
    # and back: what this mirror saves, torch.optim.Adam loads
    sd = m.optimizers[1].state_dict()
    ps = [torch.nn.Parameter(p.detach().cpu().clone()) for p in m.optimizers[1].param_groups[0]['params']]
    topt = torch.optim.Adam(ps, lr=args.lr_human, weight_decay=args.wd_human)
    topt.load_state_dict({'state': {i: {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in st.items()}
                                    for i, st in sd['state'].items()}, 'param_groups': sd['param_groups']})
    assert len(topt.state_dict()['state']) == len(ps)
    assert int(topt.state_dict()['state'][8]['step']) == 3 and int(topt.state_dict()['state'][0]['step']) == 6


def test_warmup_then_fit_adam_state_matches_torch():
    """Three warm-up iterations (the reference leaves linear_out.* without gradient: no weight decay, no Adam state, step
    count stays 0, :3455-3509) followed by one fit step, against torch.optim.Adam on the oracle with the same sampled
    batches: linear_out.* must come out of the warm-up bit-identical to their initial values and take their FIRST
    Adam step (bias correction of step 1) in the fit step while every other tensor takes its fourth."""
    V, Fr, N = 4, 12, 40
    args = assets.default_args(h_dim=48, phase_rbf_dim=12, instance_code_size=4, monotonic_network_n_nodes=16,
                               batch_size=N, lr_human=1e-3)
    params = assets.make_fit_params(args, V, seed=61)
    data = assets.make_fit_data(V, Fr, seed=62)
    _near(data, V, Fr, 63, 150.0)
    orc, m = _oracle(args, params, data, 2), _model(args, data, params, 2)
    torch.manual_seed(77)
    m.warmup(3)
    torch.manual_seed(77)
    for _ in range(3):
        vi, fi = torch.randint(0, V, size=(N,)), torch.randint(0, Fr, size=(N,))
        loss = orc.warmup_loss(vi, fi)
        orc.opt_motion.zero_grad()
        orc.opt_phase.zero_grad()
        loss.backward()
        orc.opt_motion.step()
        orc.opt_phase.step()
    got = m.engine.named_params()
    for k in ('learned_motion.linear_out.weight', 'learned_motion.linear_out.bias'):
        assert np.array_equal(npy(got[k]).reshape(-1), params[k].reshape(-1).astype(np.float64)), k
    assert int(m.engine.g_motion.step_dev) == 3 and int(m.engine.g_motion.step_lag_dev) == 0
    for k, p in orc.p.items():
        if (k.startswith('learned_motion.') or k.startswith('phase_')) and k in got:
            assert_adam_param(npy(got[k]), npy(p), 'after warm-up: ' + k, lr=1e-3)
    r = np.random.RandomState(64)
    vi, fi = torch.tensor(r.randint(0, V, N)), torch.tensor(r.randint(0, Fr, N))
    orc.step(vi, fi)
    m.step(vi, fi)
    for k, p in orc.p.items():
        if p.grad is not None and k != 'learned_betas':
            assert_adam_param(npy(got[k]), npy(p), 'after the fit step: ' + k, lr=args.lr_camera if 'camera' in k else 1e-3)
    assert int(m.engine.g_motion.step_dev) == 4 and int(m.engine.g_motion.step_lag_dev) == 1


def test_opt_cam_leaves_the_training_optimiser_fresh():
    """the reference fits the cameras with a throw-away Adam (:2869-2906): afterwards opt_cameras has no state"""
    V, Fr = 4, 12
    args = assets.default_args(h_dim=48, phase_rbf_dim=12, instance_code_size=4, monotonic_network_n_nodes=16,
                               batch_size=40)
    data = assets.make_fit_data(V, Fr, seed=2)
    m = _model(args, data, assets.make_fit_params(args, V, seed=3), 2)
    m.opt_cam(4)
    g = m.engine.g_cam
    assert int(g.step_dev) == 0 and float(g.m.abs().max()) == 0.0 and float(g.v.abs().max()) == 0.0
    assert m.opt_cameras.state_dict()['state'] == {}
