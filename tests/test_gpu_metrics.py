"""GPU parity of the evaluation metrics (csrc/metrics.cu behind nemo/utils/pose_utils.py) against outputs of the
reference's own numpy functions (tests/golden/metrics.npz) and the oracle."""
import numpy as np
import pytest
import torch

from .util import gold

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def P():
    from nemo_cvpr2023_b200.nemo.utils import pose_utils
    return pose_utils


@pytest.mark.parametrize('name', ['j', 'v', 'refl'])
def test_procrustes_matches_reference(P, name):
    g = gold('metrics')
    S1, S2 = g[name + '_S1'], g[name + '_S2']
    hat = P.compute_similarity_transform_batch(S1, S2)
    assert isinstance(hat, np.ndarray) and hat.shape == S1.shape
    assert np.abs(hat - g[name + '_hat']).max() < 2e-5 * max(1.0, np.abs(g[name + '_hat']).max())
    if name == 'refl':
        return
    e = P.reconstruction_error(S1, S2, pa=True, reduction=None)
    assert np.allclose(e, g[name + '_err_pa'], rtol=2e-5)
    assert abs(P.reconstruction_error(S1, S2, pa=True) - g[name + '_err_pa'].mean()) < 2e-5 * g[name + '_err_pa'].mean()
    e = P.reconstruction_error(S1, S2, pa=False, reduction=None)
    assert np.allclose(e, g[name + '_err'], rtol=1e-5)
    assert abs(P.reconstruction_error(S1, S2, pa=False, reduction='sum') - g[name + '_err'].sum()) < 1e-5 * g[name + '_err'].sum()
    # single-item API, both orientations, with the transform
    hat0, (s, R, t) = P.compute_similarity_transform(S1[0], S2[0], return_transform=True)
    ref = g[name + '_srt0']
    assert abs(s - ref[0]) < 1e-5 * ref[0] and np.abs(R - ref[1:10].reshape(3, 3)).max() < 1e-5
    assert np.abs(t.reshape(-1) - ref[10:]).max() < 1e-4
    assert np.abs(hat0 - g[name + '_hat'][0]).max() < 2e-5 * max(1.0, np.abs(g[name + '_hat']).max())
    hat0t = P.compute_similarity_transform(S1[0].T.copy(), S2[0].T.copy())
    assert hat0t.shape == (3, S1.shape[1]) and np.abs(hat0t.T - hat0).max() < 1e-6


def test_procrustes_is_invariant_and_exact(P):
    """properties: a similarity-transformed copy aligns back with zero error; R is a proper rotation."""
    r = np.random.RandomState(3)
    S2 = r.standard_normal((16, 49, 3)).astype(np.float32)
    q = r.standard_normal((16, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z),
                  2 * (y * z - w * x), 2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).reshape(16, 3, 3)
    S1 = (0.6 * np.einsum('bij,bpj->bpi', R, S2) + r.standard_normal((16, 1, 3))).astype(np.float32)
    hat, (s, Rg, t) = P.compute_similarity_transform_batch(S1, S2, return_transform=True)
    assert np.abs(hat - S2).max() < 2e-5
    assert np.allclose(np.linalg.det(Rg), 1.0, atol=1e-5)
    assert np.allclose(np.einsum('bij,bkj->bik', Rg, Rg), np.eye(3)[None], atol=1e-5)
    assert np.allclose(s, 1 / 0.6, rtol=1e-5)
    assert P.reconstruction_error(S1, S2, pa=True) < 2e-5
    assert P.reconstruction_error(S2, S2, pa=False) == 0.0
    tt = P.compute_similarity_transform_batch(torch.tensor(S1, device='cuda'), torch.tensor(S2, device='cuda'))
    assert isinstance(tt, torch.Tensor) and tt.is_cuda and float((tt.cpu() - torch.tensor(S2)).abs().max()) < 2e-5


def test_pck_matches_reference():
    from nemo_cvpr2023_b200._lib import lib
    g = gold('metrics')
    pred, gt, size = (torch.tensor(g[k], device='cuda') for k in ('pck_pred', 'pck_gt', 'pck_size'))
    out = torch.zeros(2, device='cuda')
    lib.pck2d(pred, gt, size, pred.shape[1], pred.shape[0], out)
    assert abs(100 * float(out[0] / out[1]) - float(g['pck'])) < 1e-4


def test_eval_3d_and_2d_match_per_frame_cpu(tmp_path):
    """MultiViewModel.eval_3d / eval_2d (batched, device) == the reference's per-frame procedure restated on the CPU
    oracle: SMPL(betas=None, body_pose, global_orient=None) per frame, reconstruction_error(pa=False), f_pck."""
    import pandas as pd
    from nemo_cvpr2023_b200 import assets
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence
    from oracle import nemo_oracle as O, pose_metrics as M
    args = assets.default_args(h_dim=32, phase_rbf_dim=8, instance_code_size=3, monotonic_network_n_nodes=10, batch_size=16)
    V, F = 2, 12
    params = assets.make_fit_params(args, V, seed=5)
    seqs = SyntheticMultiViewSequence(V, F, seed=3)
    m = nmm.NemoV2(args, seqs, 'cuda', init_params=params)
    df3 = m.eval_3d(str(tmp_path), num_frames=6)
    assert list(df3.columns) == ['mpjpe-ours', 'mpvpe-ours', 'mpjpe-vibe', 'mpvpe-vibe'] and len(df3) == V
    assert list(pd.read_csv(tmp_path / 'eval_3d.csv').columns[1:]) == list(df3.columns)
    sm = O.to_torch(assets.make_smpl_assets(0), torch.float64)
    pred = m.get_preds()
    frames = [int(np.round(c / 6 * F)) for c in range(6)]
    for v in range(V):
        def mesh(bp):
            out = [O.smpl_forward(sm, body_pose=torch.tensor(x, dtype=torch.float64)[None]) for x in bp]
            return (np.stack([o.vertices[0].numpy() for o in out]), np.stack([o.joints[0, :15].numpy() for o in out]))
        vg, jg = mesh(np.asarray(seqs.sequences[v]['pose_3d_gt'])[frames][:, 3:])
        vo, jo = mesh(pred['poses'][v][frames].double().cpu().numpy())
        vv, jv = mesh(np.asarray(seqs.sequences[v]['pose'])[frames][:, 3:-1])
        ref = [1000 * M.reconstruction_error(jg, jo, pa=False), 1000 * M.reconstruction_error(vg, vo, pa=False),
               1000 * M.reconstruction_error(jg, jv, pa=False), 1000 * M.reconstruction_error(vg, vv, pa=False)]
        assert np.allclose(df3.iloc[v].values, ref, rtol=2e-5), (df3.iloc[v].values, ref)
    dfd = m.eval_3d(str(tmp_path), dynamic_only=True)
    assert (tmp_path / 'eval_3d_dynamic.csv').exists() and list(dfd.columns)[:2] == ['mpjpe-ours', 'mpvpe-ours']
    # 2-D: ReconError2D (rmse keypoint loss) and PCK of the re-projected joints
    df2 = m.eval_2d(str(tmp_path), num_frames=6)
    assert list(df2.columns)[:4] == ['recon_error_2d-ours', 'pck-ours', 'recon_error_2d-op', 'pck-op'] and 'pck-vibe' in df2
    gt_all = m.points2d_gt_all.double().cpu().numpy()
    p2d = m.learned_camera_projection(pred['j'].reshape(-1, 25, 3), pred['view_idx'].reshape(-1)).reshape(V, F, 25, 2)
    p2d = p2d.detach().double().cpu().numpy()
    size = m.gt_bbox_size.double().cpu().numpy()
    for v in range(V):
        gt, pr = gt_all[v][frames][:, :15], p2d[v][frames][:, :15]
        rm = ((gt[..., 2:] > 0.5) * np.sqrt(1e-6 + ((pr - gt[..., :2]) ** 2).sum(-1, keepdims=True))).mean()
        assert abs(df2['recon_error_2d-ours'][v] - rm) <= 1e-5 * rm
        assert abs(df2['pck-ours'][v] - M.pck(pr, gt, size[v][frames])) < 1e-3
        assert df2['pck-op'][v] == 100.0   # the synthetic OpenPose track IS the gt track
