"""Data-parallel sharding of one fit across the GPUs of a box; host-side logic only.

The reference is single-GPU (`scripts/learned_multi_view_recon_nn.py:134`).  The fit shards naturally by video
instance (view): cameras (V,9), instance codes (V,C) and phase networks (per view) belong to exactly one rank and
never communicate; only the shared parameters -- the motion MLP and the RBF widths, one flat gradient buffer --
are summed across ranks (one all-reduce per step), after which every rank applies the identical Adam update.

Loss normalisation: every term of `step` is a mean over samples, or (2D loss) a mean over the views present of
per-view means (`nemo/neural_motion_model.py:3551-3558`).  With equal shards (same number of views and of
samples per rank) the global loss is the plain mean of the ranks' local losses, so each rank back-propagates
`local_loss / world` and the summed shared gradient equals the single-process gradient up to fp32 summation
order; per-view gradients come out already scaled for the global loss.
"""
from __future__ import annotations

import re

import numpy as np
import torch

SHARED_PREFIXES = ('learned_motion.', 'phase_rbf.')
_PHASE = re.compile(r'^phase_networks\.(\d+)\.(shifts|scales)$')


def shard_views(num_views: int, world: int, rank: int) -> list:
    """Contiguous, equal slice of the views for `rank`; the fit needs num_views % world == 0 (equal shards)."""
    if num_views % world:
        raise ValueError('data-parallel fit needs the views (%d) to divide evenly over the ranks (%d)' % (num_views, world))
    per = num_views // world
    return list(range(rank * per, (rank + 1) * per))


def shard_plan(num_views: int, num_frames: int, world: int, rank: int) -> dict:
    """Which (views, frame window) `rank` fits.  Views are split first; with more ranks than views
    (V < G) every view's frames are cut into world / V equal windows, and the view's camera / phase / instance-code
    gradients have to be summed over the ranks that share it (`reduce_all`).  Equal shards only."""
    if num_views % world == 0:
        return dict(views=shard_views(num_views, world, rank), frame_lo=0, frame_hi=num_frames, reduce_all=False)
    if world % num_views == 0 and num_frames % (world // num_views) == 0:
        per_view = world // num_views
        n = num_frames // per_view
        k = rank % per_view
        return dict(views=[rank // per_view], frame_lo=k * n, frame_hi=(k + 1) * n, reduce_all=True)
    raise ValueError('cannot cut %d views x %d frames into %d equal shards' % (num_views, num_frames, world))


def shard_fit_data_frames(data: dict, views: list, frame_lo: int, frame_hi: int) -> dict:
    """`shard_fit_data` + a frame window on every (V,F,...) entry."""
    out = shard_fit_data(data, views)
    return {k: (v[:, frame_lo:frame_hi] if hasattr(v, 'shape') and len(v.shape) >= 2 else v) for k, v in out.items()}


def is_shared(name: str) -> bool:
    return name.startswith(SHARED_PREFIXES)


def shard_fit_params(params: dict, views: list) -> dict:
    """Slice a full parameter dict (reference state_dict names) down to the given views; shared entries are kept."""
    out = {}
    for k, v in params.items():
        m = _PHASE.match(k)
        if m:
            if int(m.group(1)) in views:
                out['phase_networks.%d.%s' % (views.index(int(m.group(1))), m.group(2))] = v
        elif k in ('learned_cameras', 'learned_instance_code'):
            out[k] = v[np.asarray(views)] if isinstance(v, np.ndarray) else v[torch.as_tensor(views)]
        else:
            out[k] = v
    return out


def shard_fit_data(data: dict, views: list) -> dict:
    """Slice the per-view targets (V,F,...) down to the given views."""
    idx = np.asarray(views)
    out = {}
    for k, v in data.items():
        if hasattr(v, 'shape') and len(v.shape) >= 2:
            out[k] = v[idx] if isinstance(v, np.ndarray) else v[torch.as_tensor(views)]
        else:
            out[k] = v
    return out


def allreduce_shared_(grads, group=None):
    """Sum the shared-parameter gradients over the ranks in ONE collective: a flat tensor is reduced in place,
    a list of tensors is flattened, reduced and scattered back."""
    import torch.distributed as dist
    if isinstance(grads, torch.Tensor):
        dist.all_reduce(grads, group=group)
        return grads
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()
    return grads
