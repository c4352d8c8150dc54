"""Shared helpers for the parity tests: golden loading and the tolerance contract."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# BASELINE.json north_star tolerances
TOL_VERT_ABS = 1e-5     # vertices & joints, metres, absolute
TOL_LOSS_REL = 1e-5     # losses, relative
TOL_GRAD_REL = 1e-4     # gradients, relative (per-tensor L2 and max-norm)


def gold(name):
    return dict(np.load(os.path.join(GOLD, name + '.npz')))


def T(a, grad=False, dtype=torch.float32, device='cpu'):
    a = np.asarray(a)
    if a.dtype.kind in 'iu':
        return torch.tensor(a, dtype=torch.long, device=device)
    t = torch.tensor(a, dtype=dtype, device=device)
    return t.requires_grad_(True) if grad else t


def npy(t):
    return t.detach().cpu().double().numpy()


def grad_err(g, ref):
    """(relative L2 error, max-abs error relative to max|ref|) of a gradient tensor."""
    g, ref = np.asarray(g, np.float64), np.asarray(ref, np.float64)
    n = np.linalg.norm(ref)
    l2 = np.linalg.norm(g - ref) / max(n, 1e-30)
    mx = np.abs(g - ref).max() / max(np.abs(ref).max(), 1e-30)
    return l2, mx


def assert_grad(g, ref, tol=TOL_GRAD_REL, what='', zero_atol=1e-3):
    if np.abs(np.asarray(ref)).max() == 0:
        # exactly cancelling gradient in the reference (e.g. d/d linear_out.bias of trans - trans_0):
        # relative error is undefined, only summation-order noise may remain
        assert np.abs(np.asarray(g)).max() <= zero_atol, '%s: reference gradient is 0, got %.3e' % (what, np.abs(g).max())
        return
    l2, mx = grad_err(g, ref)
    assert l2 <= tol and mx <= tol, '%s: grad rel-L2 %.3e, rel-max %.3e > %.1e' % (what, l2, mx, tol)


def assert_abs(a, ref, tol, what=''):
    e = np.abs(np.asarray(a, np.float64) - np.asarray(ref, np.float64)).max()
    assert e <= tol, '%s: max abs err %.3e > %.1e' % (what, e, tol)


def assert_rel(a, ref, tol, what=''):
    a, ref = float(a), float(ref)
    assert abs(a - ref) <= tol * max(abs(ref), 1e-30), '%s: %r vs %r (rel %.3e > %.1e)' % (
        what, a, ref, abs(a - ref) / max(abs(ref), 1e-30), tol)


def assert_adam_param(p, ref, what='', lr=1e-1):
    """Parameters after Adam steps.  Adam moves every weight by ~lr * g/(|g| + eps): entries whose gradient is
    summation-order noise around 0 can legitimately differ by a fraction of lr, so the bulk must agree
    tightly (99.5% within 2e-5 of the value scale) and no entry may be off by more than one lr-sized step."""
    p, ref = np.asarray(p, np.float64), np.asarray(ref, np.float64)
    d = np.abs(p - ref)
    scale = max(1.0, np.abs(ref).max())
    frac = (d > 2e-5 * scale).mean()
    assert frac <= 5e-3, '%s: %.3f%% of entries differ by more than 2e-5' % (what, 100 * frac)
    assert d.max() <= 2.5 * lr, '%s: max abs err %.3e' % (what, d.max())


def make_demo_tree(root, seed=11):
    """A seeded stand-in for a custom-video experiment directory in the layout `DemoMultiViewSequence` reads
    (nemo/multi_view_sequence.py:92-222): two videos with different frame counts and image sizes, OpenPose BODY_25
    JSONs (one frame without a detection), a VIBE pickle with two tracked people (one present on a subset of frames).
    Returns the `nemo_data_cfg` dict."""
    import json

    import cv2
    import joblib
    r = np.random.RandomState(seed)
    names = ['swing.0', 'swing.1']
    for vi, (name, T, hw) in enumerate(zip(names, (14, 11), ((48, 64), (40, 72)))):
        fr, op, vb = (os.path.join(root, name + s) for s in ('.frames', '.op', '.vibe'))
        for d in (fr, op, vb):
            os.makedirs(d, exist_ok=True)
        centre = r.uniform(10, 30, (T, 1, 2))
        for t in range(T):
            cv2.imwrite(os.path.join(fr, f"{t + 1:06d}.png"), r.randint(0, 255, hw + (3,)).astype(np.uint8))
            kp = np.concatenate([centre[t] + r.standard_normal((25, 2)), r.uniform(0, 1, (25, 1))], 1)
            people = [] if (vi == 0 and t == 3) else [{'pose_keypoints_2d': kp.reshape(-1).tolist()}]
            json.dump({'people': people}, open(os.path.join(op, f"{t + 1:06d}_keypoints.json"), 'w'))

        def person(ids, offset):
            n = len(ids)
            return {'frame_ids': np.asarray(ids), 'betas': r.standard_normal((n, 10)).astype(np.float32),
                    'pose': r.standard_normal((n, 72)).astype(np.float32),
                    'orig_cam': r.standard_normal((n, 4)).astype(np.float32),
                    'verts': r.standard_normal((n, 20, 3)).astype(np.float32),
                    'joints2d_img_coord': (2 * centre[ids] + offset + r.standard_normal((n, 49, 2))).astype(np.float32),
                    'joints3d': None}
        joblib.dump({1: person(list(range(2, T - 1)), 40.0), 2: person(list(range(T)), 1.0)},
                    os.path.join(vb, 'vibe_output.pkl'))
    joblib.dump({'R': np.eye(3)}, os.path.join(root, names[0] + '.cam.pickle'))
    return {'exp_dir': root, 'videos': {'names': names}}
