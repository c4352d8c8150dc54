"""Generates tests/golden/metrics.npz from the UNMODIFIED reference functions (numpy):
/root/reference/nemo/utils/pose_utils.py compute_similarity_transform / reconstruction_error, and the f_pck closure
of eval_2d (nemo/neural_motion_model.py:524-532, restated with torch ops exactly as written there since a closure
cannot be imported).  Run in the build container only:  python -m oracle.make_golden_metrics"""
import os
import sys

import numpy as np
import torch

from . import ref_harness as H


def main():
    H._install_stubs()
    sys.path.insert(0, H.REF)
    from nemo.utils import pose_utils as P
    r = np.random.RandomState(7)
    out = {}
    for name, (B, Pn) in {'j': (6, 15), 'v': (2, 6890)}.items():
        S2 = r.standard_normal((B, Pn, 3)).astype(np.float32)
        A = r.standard_normal((B, 3, 3))
        Rm = np.stack([np.linalg.qr(a)[0] for a in A])
        Rm[np.linalg.det(Rm) < 0, :, 0] *= -1
        S1 = (1.7 * np.einsum('bij,bpj->bpi', Rm, S2) + r.standard_normal((B, 1, 3))
              + 0.05 * r.standard_normal((B, Pn, 3))).astype(np.float32)
        out[name + '_S1'], out[name + '_S2'] = S1, S2
        out[name + '_hat'] = P.compute_similarity_transform_batch(S1.astype(np.float64), S2.astype(np.float64)).astype(np.float32)
        out[name + '_err_pa'] = P.reconstruction_error(S1.astype(np.float64), S2.astype(np.float64), pa=True, reduction=None)
        out[name + '_err'] = P.reconstruction_error(S1.astype(np.float64), S2.astype(np.float64), pa=False, reduction=None)
        hat0, (s, R, t) = P.compute_similarity_transform(S1[0].astype(np.float64), S2[0].astype(np.float64), return_transform=True)
        out[name + '_srt0'] = np.concatenate([[s], R.reshape(-1), t.reshape(-1)])
    # a reflected pair: the optimal orthogonal map is improper, the reference forces det(R) = +1
    S2 = r.standard_normal((1, 40, 3)).astype(np.float32)
    S1 = (S2 * np.array([1, 1, -1], np.float32) + 0.01 * r.standard_normal((1, 40, 3))).astype(np.float32)
    out['refl_S1'], out['refl_S2'] = S1, S2
    out['refl_hat'] = P.compute_similarity_transform_batch(S1.astype(np.float64), S2.astype(np.float64))
    # PCK closure (:524-532), torch ops as in the reference
    B, J = 9, 15
    pred = torch.tensor(r.uniform(0, 1000, (B, J, 2)), dtype=torch.float32)
    gt = torch.cat([pred + torch.tensor(r.standard_normal((B, J, 2)) * 20, dtype=torch.float32),
                    torch.tensor(r.uniform(0, 1, (B, J, 1)), dtype=torch.float32)], -1)
    gt_size = torch.tensor(r.uniform(200, 600, (B,)), dtype=torch.float32)
    crit = torch.nn.MSELoss(reduction='none')
    rmse = torch.sqrt(1e-6 + crit(pred, gt[..., :2]).sum(-1, keepdim=True))
    mask = (gt[..., 2:] > 0.5).float()
    count = (mask * (rmse < (0.05 * gt_size.unsqueeze(-1).unsqueeze(-1))).float()).sum()
    out['pck_pred'], out['pck_gt'], out['pck_size'] = pred.numpy(), gt.numpy(), gt_size.numpy()
    out['pck'] = np.float64(100 * count / mask.sum())
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'metrics.npz')
    np.savez_compressed(dst, **out)
    print('wrote', dst, {k: np.shape(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
