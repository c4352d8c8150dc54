"""Generate tests/golden/ref_ckpt_v3.pt + ref_ckpt_v3.npz: a checkpoint WRITTEN BY THE UNMODIFIED REFERENCE
(`MultiViewModel.save`, nemo/neural_motion_model.py:257-265) after its own warm-up and two fit steps, plus what the
reference itself computes after loading it back (`load` :267-280) and taking one more step.  Build container only.

    python -m oracle.make_golden_ckpt

The warm-up matters: the reference's warm-up leaves `linear_out.*` without gradient, so torch.optim.Adam keeps no state
for them and their per-parameter `step` lags the other tensors of the same optimiser (:3455-3509) -- the layout a
loader of reference checkpoints has to reproduce.  The `smpl.*`, `vp.*`, `pose_prior.*` entries of `model_sd` (20 MB of
frozen asset buffers that the reference's `load` itself drops, :269-275) are removed from the committed file.
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
from oracle.make_golden import GOLD, npy, step_case  # noqa: E402

DROP = ('vp', 'pose_prior', 'renderer', 'smpl')


def main():
    torch.set_num_threads(1)
    sm, vp_np, gmm_np = assets.make_smpl_assets(0), assets.make_vposer_weights(2), assets.make_gmm(3)
    mods = rh.setup(sm)
    args, params, data, view_idx, frame_idx = step_case(3)
    vi, fi = torch.tensor(view_idx), torch.tensor(frame_idx)
    m = rh.build_fit_model(mods, args, params, sm, vp_np, gmm_np, data, version=3)
    torch.manual_seed(123)
    m.warmup(3)
    for _ in range(2):
        m.step(vi, fi)
    path = os.path.join(GOLD, 'ref_ckpt_v3.pt')
    m.save(path)
    saved = torch.load(path, weights_only=False)
    saved['model_sd'] = {k: v for k, v in saved['model_sd'].items() if not k.startswith(DROP)}
    torch.save(saved, path)
    # what the reference does with its own checkpoint: a fresh model, load, one step
    m2 = rh.build_fit_model(mods, args, assets.make_fit_params(args, 3, seed=999), sm, vp_np, gmm_np, data, version=3)
    m2.load(path)
    out = {}
    for k, p in m2.named_parameters():
        if p.requires_grad:
            out['loaded_param.' + k] = npy(p)
    loss_dict, _ = m2.step(vi, fi)
    for k, v in loss_dict.items():
        out['loss.' + k] = np.asarray(v, dtype=np.float64)
    for k, p in m2.named_parameters():
        if p.requires_grad:
            out['after_param.' + k] = npy(p)
    steps = [[int(s['step']) for s in o.state_dict()['state'].values()] for o in m2.optimizers]
    out['opt_steps_motion'] = np.asarray(steps[1])
    np.savez_compressed(os.path.join(GOLD, 'ref_ckpt_v3.npz'), **out)
    os.chdir(ROOT)
    print('wrote', path, os.path.getsize(path) // 1024, 'KB; optimiser steps after the extra step:', steps)


if __name__ == '__main__':
    main()
