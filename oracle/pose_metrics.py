"""ORACLE (test infrastructure only): CPU restatement of the reference's evaluation metrics.

    compute_similarity_transform   nemo/utils/pose_utils.py:78-137
    reconstruction_error           nemo/utils/pose_utils.py:148-159
    pck                            nemo/neural_motion_model.py:524-532 (closure f_pck of eval_2d)

Pinned against outputs of the reference's own functions: `oracle/make_golden_metrics.py` imports
/root/reference/nemo/utils/pose_utils.py (numpy) and writes tests/golden/metrics.npz;
tests/test_oracle_golden.py::test_metrics_oracle checks this file against it.  Only tests/ may import this module."""
import numpy as np


def compute_similarity_transform(S1, S2):
    """S1, S2 (P,3) float64 -> (S1_hat (P,3), scale, R (3,3), t (3,))   [:78-137, transposed convention folded in]"""
    X1, X2 = S1.T.astype(np.float64), S2.T.astype(np.float64)
    mu1, mu2 = X1.mean(1, keepdims=True), X2.mean(1, keepdims=True)
    A, B = X1 - mu1, X2 - mu2
    var1 = (A ** 2).sum()
    K = A @ B.T
    U, _, Vh = np.linalg.svd(K)
    V = Vh.T
    Z = np.eye(3)
    Z[-1, -1] *= np.sign(np.linalg.det(U @ V.T))
    R = V @ Z @ U.T
    scale = np.trace(R @ K) / var1
    t = mu2 - scale * (R @ mu1)
    return (scale * R @ X1 + t).T, scale, R, t[:, 0]


def reconstruction_error(S1, S2, pa=True, reduction='mean'):
    """S1, S2 (B,P,3)   [:148-159]"""
    hat = np.stack([compute_similarity_transform(a, b)[0] for a, b in zip(S1, S2)]) if pa else S1
    re = np.sqrt(((hat - S2) ** 2).sum(-1)).mean(-1)
    return re.mean() if reduction == 'mean' else (re.sum() if reduction == 'sum' else re)


def pck(pred, gt, gt_size):
    """pred (B,J,2), gt (B,J,3), gt_size (B,) -> percentage   [neural_motion_model.py:524-532]"""
    rmse = np.sqrt(1e-6 + ((pred - gt[..., :2]) ** 2).sum(-1))
    mask = (gt[..., 2] > 0.5).astype(np.float64)
    return 100.0 * (mask * (rmse < 0.05 * gt_size[:, None])).sum() / mask.sum()
