"""Mirror of `nemo/utils/pose_utils.py:78-159` (Procrustes alignment and reconstruction error) on the device.

The reference loops over frames in numpy (`compute_similarity_transform_batch`); here the whole batch is one launch
of `nemo_b200_procrustes` / `nemo_b200_point_error` (csrc/metrics.cu).  Inputs may be numpy arrays or tensors;
results come back in the type the reference returns (numpy)."""
from __future__ import annotations

import numpy as np
import torch

from ..._lib import lib, require_cuda


def _dev(x):
    require_cuda()
    if isinstance(x, torch.Tensor):
        return x.detach().to('cuda', torch.float32).contiguous()
    return torch.as_tensor(np.asarray(x), dtype=torch.float32, device='cuda').contiguous()


def compute_similarity_transform_batch(S1, S2, return_transform=False):
    """:140-145.  S1, S2 (B,P,3) -> S1_hat (B,P,3) [, (scale (B,), R (B,3,3), t (B,3))]."""
    a, b = _dev(S1), _dev(S2)
    assert a.shape == b.shape and a.dim() == 3 and a.shape[-1] == 3
    B, P = a.shape[:2]
    out = torch.empty_like(a)
    srt = torch.empty(B, 13, device=a.device)
    lib.procrustes(a, b, P, B, out, None, srt, )
    if isinstance(S1, torch.Tensor):
        res, cv = out.to(S1.device), (lambda x: x.to(S1.device))
    else:
        res, cv = out.cpu().numpy(), (lambda x: x.cpu().numpy())
    if return_transform:
        return res, (cv(srt[:, 0]), cv(srt[:, 1:10].reshape(B, 3, 3)), cv(srt[:, 10:13]))
    return res


def compute_similarity_transform(S1, S2, return_transform=False):
    """:78-137.  Single item, (3,P) or (P,3) (the reference transposes when the first dimension is not 3 or 2)."""
    S1n = S1 if isinstance(S1, torch.Tensor) else np.asarray(S1)
    transposed = S1n.shape[0] not in (3, 2)
    a, b = _dev(S1), _dev(S2)
    if not transposed:
        a, b = a.t().contiguous(), b.t().contiguous()
    assert a.shape == b.shape and a.shape[1] == 3, 'compute_similarity_transform takes 3-D point sets'
    hat, (s, R, t) = compute_similarity_transform_batch(a[None], b[None], return_transform=True)
    hat = hat[0] if transposed else hat[0].t()
    hat = hat.to(S1.device) if isinstance(S1, torch.Tensor) else hat.cpu().numpy()
    if return_transform:
        s, R, t = s[0], R[0], t[0].reshape(3, 1)
        if not isinstance(S1, torch.Tensor):
            s, R, t = float(s), R.cpu().numpy(), t.cpu().numpy()
        return hat, (s, R, t)
    return hat


def reconstruction_error(S1, S2, pa=True, reduction='mean'):
    """:148-159.  Mean per-point distance after (pa=True) Procrustes alignment of S1 onto S2; numpy scalar / (B,)."""
    a, b = _dev(S1), _dev(S2)
    B, P = a.shape[:2]
    err = torch.empty(B, device=a.device)
    if pa:
        lib.procrustes(a, b, P, B, None, err, None)
    else:
        lib.point_error(a, b, P, B, err)
    re = err.double().cpu().numpy()
    if reduction == 'mean':
        return re.mean()
    if reduction == 'sum':
        return re.sum()
    return re
