"""GPU parity of the pre-phases and checkpointing of the fit:
`warmup` (nemo/neural_motion_model.py:3455-3509), `opt_cam` (:2869-2906) and `save` / `load` (:257-280), against the
CPU oracle restatement of the same reference lines (autograd + torch.optim.Adam)."""
import numpy as np
import pytest
import torch

from nemo_cvpr2023_b200 import assets
from oracle import nemo_oracle as O
from tests.util import TOL_LOSS_REL, assert_grad, assert_rel, npy

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _pair(version=2, V=4, Fr=12, seed=21, **over):
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence
    args = assets.default_args(h_dim=48, phase_rbf_dim=12, instance_code_size=4, monotonic_network_n_nodes=16,
                               batch_size=40, **over)
    params = assets.make_fit_params(args, V, seed=seed)
    data = assets.make_fit_data(V, Fr, seed=seed + 1)
    r = np.random.RandomState(seed)
    data['points2d_gt_all'][..., 0] = 540 + 150 * r.standard_normal((V, Fr, 25))
    data['points2d_gt_all'][..., 1] = 960 + 150 * r.standard_normal((V, Fr, 25))
    orc = O.FitOracle(args, O.to_torch(params), O.to_torch(assets.make_smpl_assets(0)),
                      O.to_torch(assets.make_vposer_weights(2)), O.to_torch(assets.make_gmm(3)), O.to_torch(data),
                      version=version).build_optimizers()
    cls = {2: nmm.NemoV2, 3: nmm.NemoV3}[version]
    m = cls(args, SyntheticMultiViewSequence(data=data), DEV, init_params=params)
    return args, orc, m, r


def test_warmup_step_matches_oracle():
    """one warm-up iteration on a fixed batch: loss and every gradient the reference's `loss.backward()` produces"""
    args, orc, m, r = _pair()
    vi, fi = torch.tensor(r.randint(0, 4, 40)), torch.tensor(r.randint(0, 12, 40))
    loss = orc.warmup_loss(vi, fi)
    loss.backward()
    w, _ = m.engine.forward_backward(vi.to(DEV), fi.to(DEV), update=True, mode='warmup')
    assert_rel(float(w['acc'][2]), float(loss), TOL_LOSS_REL, 'warmup loss')
    grads = m.engine.named_grads()
    n = 0
    for k, p in orc.p.items():
        if p.grad is not None and (k.startswith('learned_motion.') or k.startswith('phase_') or k.startswith('instance')):
            if float(p.grad.abs().max()) == 0.0:
                assert float(grads[k].abs().max()) == 0.0, k
                continue
            assert_grad(npy(grads[k]), npy(p.grad), what='warmup ' + k)
            n += 1
    assert n >= 8


def test_warmup_loop_decreases_the_pose_loss():
    args, orc, m, r = _pair(lr_human=1e-3)
    torch.manual_seed(0)
    losses = m.warmup(30)
    assert len(losses) == 30 and all(np.isfinite(losses))
    assert np.mean(losses[-5:]) < np.mean(losses[:5])
    assert m.warmup(0) == []


def test_opt_cam_matches_oracle():
    """MultiViewModel.opt_cam: a fresh Adam on the cameras, first frame of every view, plain mean of the keypoint loss"""
    args, orc, m, r = _pair()
    V = 4
    cam = orc.p['learned_cameras']
    opt = torch.optim.Adam([cam], lr=args.lr_camera)
    vi, fi = torch.arange(V), torch.zeros(V, dtype=torch.long)
    ref_log = []
    for _ in range(5):
        opt.zero_grad()
        pred = orc.get_preds_batch(vi, fi)
        p2d = O.learned_camera_projection(pred['j'], vi, cam, orc.focal, orc.data['IMG_D0'], orc.data['IMG_D1'])
        gt = orc.data['points2d_gt_all'][vi, fi]
        loss = O.keypoint_loss(p2d, gt[..., :2], gt[..., 2:], orc.data['gt_bbox_size'][vi, fi],
                               loss_type=args.loss).mean()
        loss.backward()
        ref_log.append(float(loss))
        opt.step()
    log = m.opt_cam(5)
    assert len(log) == 5
    for a, b in zip(log, ref_log):
        assert_rel(float(a), b, 1e-4, 'opt_cam loss')   # 5 Adam steps with lr 0.1 amplify fp32 rounding
    got = npy(m.engine.named_params()['learned_cameras'])
    assert np.abs(got - npy(cam)).max() < 2e-3 * max(1.0, np.abs(npy(cam)).max())


def test_save_load_roundtrip(tmp_path):
    """state_dict keys are the reference's (`learned_motion.*`, `phase_networks.N.*`, `learned_cameras`, ...); a loaded
    model reproduces the saved one's next step (up to the summation order of the atomics / bulk reductions)."""
    args, orc, m, r = _pair(version=3, weight_instance_loss=0.1, weight_3d_loss=5.0)
    vi, fi = torch.tensor(r.randint(0, 4, 40)), torch.tensor(r.randint(0, 12, 40))
    m.step(vi, fi)
    path = str(tmp_path / 'ckpt.pt')
    m.save(path)
    sd = torch.load(path, weights_only=False)
    keys = set(sd['model_sd'])
    for k in orc.p:
        if k != 'phase_rbf.centres':
            assert k in keys, k
    assert len(sd['opt_sd']) == len(m.optimizers)
    m2 = type(m)(args, m.multi_view_seqs, DEV, init_params=assets.make_fit_params(args, 4, seed=99))
    m2.load(path)
    a, _ = m.step(vi, fi)
    b, _ = m2.step(vi, fi)
    for k in a:
        assert abs(float(a[k]) - float(b[k])) <= 1e-6 * abs(float(a[k])) + 1e-12, (k, a[k], b[k])
    pa, pb = m.engine.named_params(), m2.engine.named_params()
    for k in pa:
        assert torch.allclose(pa[k], pb[k], rtol=1e-4, atol=1e-6), k


def test_v4_opt_cam_iteration_matches_oracle():
    """NemoV4.opt_cam (:4060-4149): per-view-mean 2D loss with the articulated pose detached + weighted 3D pose loss;
    every optimiser steps.  One iteration on a fixed batch: all gradients."""
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence
    V, Fr = 4, 12
    args = assets.default_args(h_dim=48, phase_rbf_dim=12, instance_code_size=4, monotonic_network_n_nodes=16,
                               batch_size=40, weight_3d_loss=7.0)
    params = assets.make_fit_params(args, V, seed=31)
    data = assets.make_fit_data(V, Fr, seed=32)
    r = np.random.RandomState(33)
    data['points2d_gt_all'][..., 0] = 540 + 150 * r.standard_normal((V, Fr, 25))
    data['points2d_gt_all'][..., 1] = 960 + 150 * r.standard_normal((V, Fr, 25))
    orc = O.FitOracle(args, O.to_torch(params), O.to_torch(assets.make_smpl_assets(0)),
                      O.to_torch(assets.make_vposer_weights(2)), O.to_torch(assets.make_gmm(3)), O.to_torch(data),
                      version=4).build_optimizers()
    m = nmm.NemoV4(args, SyntheticMultiViewSequence(data=data), DEV, init_params=params)
    vi, fi = torch.tensor(r.randint(0, V, 40)), torch.tensor(r.randint(0, Fr, 40))
    pred = orc.get_preds_batch(vi, fi, detach_pose=True)
    p2d = O.learned_camera_projection(pred['j'], vi, orc.p['learned_cameras'], orc.focal, orc.data['IMG_D0'],
                                      orc.data['IMG_D1'])
    gt = orc.data['points2d_gt_all'][vi, fi]
    loss_all = O.keypoint_loss(p2d, gt[..., :2], gt[..., 2:], orc.data['gt_bbox_size'][vi, fi], args.loss)
    loss = O.per_view_mean(loss_all, gt[..., 2:], vi)
    l3 = O.keypoint_loss(pred['poses'], orc.data['hmr_theta'][vi, fi], orc.data['hmr_mask'][vi, fi],
                         loss_type='mse_robust').mean()
    (loss + args.weight_3d_loss * l3).backward()
    m.engine.forward_backward(vi.to(DEV), fi.to(DEV), update=True, mode='opt_cam', detach_pose=True)
    grads = m.engine.named_grads()
    n = 0
    for k, p in orc.p.items():
        if p.grad is None or k == 'learned_betas':
            continue
        if float(p.grad.abs().max()) == 0.0:
            assert float(grads[k].abs().max()) == 0.0, k
            continue
        assert_grad(npy(grads[k]), npy(p.grad), what='V4 opt_cam ' + k)
        n += 1
    assert n >= 10
    torch.manual_seed(1)
    m.opt_cam(3)   # the loop itself: samples, detaches, steps every optimiser
    assert all(torch.isfinite(p).all() for p in m.engine.named_params().values())


def test_get_preds_given_phases_and_view_matches_get_preds_batch():
    """:3733-3781: the motion field at ALREADY WARPED phases.  Feeding it the warped phases of a batch must reproduce
    `get_preds_batch` (fused engine) through the autograd ops: vertices, the 25 loss joints out of the 49, poses,
    orientation, translation -- and it is differentiable like the reference's."""
    args, orc, m, r = _pair()
    vi, fi = torch.tensor(r.randint(0, 4, 40)), torch.tensor(r.randint(0, 12, 40))
    ref = m.get_preds_batch(vi, fi)
    warped = m.engine.ws(40)['warped'].clone()
    v, j, poses, orient, orient_aa, trans = m.get_preds_given_phases_and_view(vi, warped.view(-1, 1))
    assert v.shape == (40, 6890, 3) and j.shape == (40, 49, 3)
    idx = [38] + list(range(1, 25))
    assert float((v - ref['v']).abs().max()) < 2e-5 and float((j[:, idx] - ref['j']).abs().max()) < 2e-5
    assert float((poses - ref['poses']).abs().max()) < 1e-5 and float((orient - ref['orient']).abs().max()) < 1e-5
    assert float((trans - ref['trans']).abs().max()) < 1e-5 and float((orient_aa - ref['orient_aa']).abs().max()) < 1e-5
    o = orc.get_preds_batch(vi, fi)
    assert float((v.cpu() - o['v']).abs().max()) < 2e-5
    ph = warped.view(-1, 1).clone().requires_grad_(True)
    m.get_preds_given_phases_and_view(vi, ph)[1].sum().backward()
    assert torch.isfinite(ph.grad).all() and float(ph.grad.abs().max()) > 0


def test_step_with_pending_losses_equals_the_blocking_step():
    """`step(..., sync=False)` queues the loss read-back into pinned memory (`PendingLosses`): read late, a step's
    losses are the ones the blocking call returns (up to the summation order of the atomics), and a result that is
    still held when its pinned slot comes round again (8 in rotation) keeps its values"""
    from nemo_cvpr2023_b200.fit_engine import PendingLosses
    _, _, ma, r = _pair()
    _, _, mb, _ = _pair()
    vi, fi = torch.tensor(r.randint(0, 4, 40)), torch.tensor(r.randint(0, 12, 40))
    first_blocking = ma.step(vi, fi)[0]
    pending = []
    for _ in range(12):
        out, info = mb.step(vi, fi, sync=False)
        assert isinstance(out, PendingLosses)
        pending.append(out)
    res = [p.result() for p in pending]
    assert all(p.done() for p in pending)
    assert set(res[0]) == set(first_blocking)
    for k in first_blocking:
        assert_rel(float(res[0][k]), float(first_blocking[k]), TOL_LOSS_REL, 'first step ' + k)
    totals = [float(x['total_loss']) for x in res]
    assert all(np.isfinite(totals)) and len(set(totals)) > 1     # every step has its own values
    # no update: the same forward read both ways
    a = mb.step(vi, fi, update=False)[0]
    b = mb.step(vi, fi, update=False, sync=False)[0].result()
    for k in a:
        assert_rel(float(b[k]), float(a[k]), TOL_LOSS_REL, 'no update ' + k)
