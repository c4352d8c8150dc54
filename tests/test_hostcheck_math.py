"""CPU pin of the per-element math the CUDA kernels execute (rotmath.cuh / pathmath.cuh are
host+device headers; tests/hostcheck compiles them with g++).  Compared against the golden fixtures
(reference outputs) and the oracle.  No GPU needed; the product never loads this library."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from nemo_cvpr2023_b200 import assets
from oracle import nemo_oracle as O
from tests.util import T, assert_abs, assert_grad, gold, npy

HC_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hostcheck')


@pytest.fixture(scope='module')
def hc():
    so = os.path.join(HC_DIR, 'libhostcheck.so')
    src = os.path.join(HC_DIR, 'hostcheck.cpp')
    hdrs = [os.path.join(HC_DIR, '..', '..', 'nemo_cvpr2023_b200', 'csrc', f) for f in ('rotmath.cuh', 'pathmath.cuh', 'common.cuh')]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', '-x', 'c++', '-I/usr/local/cuda/include',
                               '-o', so, src])
    return ctypes.CDLL(so)


def fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def test_rot6d(hc):
    g = gold('geometry')
    x = f32(g['rot6d_in'])
    R = np.zeros((len(x), 9), np.float32)
    hc.hc_rot6d(fp(x), fp(R), len(x))
    assert_abs(R.reshape(-1, 3, 3), g['rot6d_out'], 1e-6)
    dx = np.zeros_like(x)
    hc.hc_rot6d_bwd(fp(x), fp(f32(g['rot6d_w'])), fp(dx), len(x))
    assert_grad(dx, g['rot6d_grad'])


@pytest.mark.parametrize('kind', [0, 1])
def test_rodrigues(hc, kind):
    g = gold('geometry')
    key = 'rodq' if kind == 0 else 'rods'
    aa = f32(g['aa_in'])
    R = np.zeros((len(aa), 9), np.float32)
    hc.hc_rod(fp(aa), fp(R), len(aa), kind)
    assert_abs(R.reshape(-1, 3, 3), g[key + '_out'], 1e-6)
    d = np.zeros_like(aa)
    hc.hc_rod_bwd(fp(aa), fp(f32(g['rod_w'])), fp(d), len(aa), kind)
    assert_grad(d, g[key + '_grad'])


@pytest.mark.parametrize('nz', [1, 0])
def test_rotmat2aa(hc, nz):
    g = gold('geometry')
    key = 'r2aa_hmr' if nz else 'r2aa_tgm'
    R = f32(g['r2aa_in'])
    aa = np.zeros((len(R), 3), np.float32)
    hc.hc_r2aa(fp(R), fp(aa), len(R), nz)
    assert_abs(aa, g[key], 2e-6)
    dR = np.zeros_like(R)
    hc.hc_r2aa_bwd(fp(R), fp(f32(g['r2aa_w'])), fp(dR), len(R), nz)
    assert_grad(dR, g[key + '_grad'])
    # the identity singularity: finite, aa == 0, no NaN gradient (the reference's autograd gives NaN there)
    eye = f32(np.eye(3)[None])
    a0, d0 = np.ones((1, 3), np.float32), np.zeros((1, 9), np.float32)
    hc.hc_r2aa(fp(eye), fp(a0), 1, nz)
    hc.hc_r2aa_bwd(fp(eye), fp(f32(np.ones((1, 3)))), fp(d0), 1, nz)
    assert np.all(a0 == 0) and np.all(np.isfinite(d0))


def test_chain_vs_oracle(hc):
    sm = O.to_torch(assets.make_smpl_assets(0))
    r = np.random.RandomState(7)
    n = 5
    rot = O.batch_rodrigues_quat(torch.tensor(0.5 * r.standard_normal((n * 24, 3)), dtype=torch.float32)).view(n, 24, 3, 3)
    J = torch.tensor(0.3 * r.standard_normal((n, 24, 3)), dtype=torch.float32)
    rot.requires_grad_(True)
    J.requires_grad_(True)
    posed, A = O.batch_rigid_transform(rot, J, sm['parents'])
    dA = torch.tensor(r.standard_normal((n, 24, 3, 4)), dtype=torch.float32)
    dJP = torch.tensor(r.standard_normal((n, 24, 3)), dtype=torch.float32)
    ((A[:, :, :3] * dA).sum() + (posed * dJP).sum()).backward()
    parents = np.asarray(assets.SMPL_PARENTS, np.int32)
    Rn, Jn = f32(npy(rot)), f32(npy(J))
    Ao, JPo = np.zeros((n, 24, 12), np.float32), np.zeros((n, 24, 3), np.float32)
    hc.hc_chain(fp(Rn), fp(Jn), 72, fp(parents), fp(Ao), fp(JPo), n)
    assert_abs(Ao.reshape(n, 24, 3, 4), npy(A)[:, :, :3], 2e-6)
    assert_abs(JPo, npy(posed), 2e-6)
    dR, dJ = np.zeros((n, 24, 9), np.float32), np.zeros((n, 24, 3), np.float32)
    hc.hc_chain_bwd(fp(Rn), fp(Jn), 72, fp(parents), fp(f32(npy(dA))), fp(f32(npy(dJP))), fp(dR), fp(dJ), n)
    assert_grad(dR.reshape(n, 24, 3, 3), npy(rot.grad))
    assert_grad(dJ, npy(J.grad))


def test_projection(hc):
    g = gold('geometry')
    for b in range(5):
        P = g['pp_pts'].shape[1]
        uv = np.zeros((P, 2), np.float32)
        R, t, p = f32(g['pp_rot'][b]), f32(g['pp_tr'][b]), f32(g['pp_pts'][b])
        hc.hc_project(fp(R), fp(t), ctypes.c_float(5000.0), ctypes.c_float(540.0), ctypes.c_float(960.0),
                      ctypes.c_float(0.0), fp(p), fp(uv), P)
        assert_abs(uv, g['pp_hmr'][b], 2e-3)
        dp, dR, dt = np.zeros((P, 3), np.float32), np.zeros(9, np.float32), np.zeros(3, np.float32)
        hc.hc_project_bwd(fp(R), fp(t), ctypes.c_float(5000.0), ctypes.c_float(540.0), ctypes.c_float(960.0),
                          ctypes.c_float(0.0), fp(p), fp(f32(g['pp_w'][b])), fp(dp), fp(dR), fp(dt), P)
        assert_grad(dp, g['pp_hmr_gpts'][b])
        assert_grad(dR.reshape(3, 3), g['pp_hmr_grot'][b])
        assert_grad(dt, g['pp_hmr_gtr'][b])
        hc.hc_project(fp(R), fp(t), ctypes.c_float(5000.0), ctypes.c_float(540.0), ctypes.c_float(960.0),
                      ctypes.c_float(1e-8), fp(p), fp(uv), P)
        assert_abs(uv, g['pp_utils'][b], 2e-3)


def test_keypoint_losses(hc):
    g = gold('modules')
    names = ['rmse', 'rmse_resized', 'mse', 'rmse_robust', 'mse_robust', 'mse_robust_resized']
    pred, gt, conf, size = f32(g['kl_pred']), f32(g['kl_gt']), f32(g['kl_conf']), f32(g['kl_size'])
    n, P = pred.shape[:2]
    for mode, nm in enumerate(names):
        ref = g['kl_' + nm]
        out = np.zeros(ref.shape, np.float32)
        dr = np.zeros_like(pred)
        hc.hc_kploss(mode, fp(pred), fp(gt), fp(conf), fp(size), n * P, P, ctypes.c_float(100.0), fp(out), fp(dr))
        assert_abs(out, ref, 1e-5 * max(1.0, np.abs(ref).max()), nm)
        assert_grad(dr, g['kl_%s_g' % nm], what=nm)


def test_rbf_bases(hc):
    names = ['gaussian', 'linear', 'quadratic', 'inverse_quadratic', 'multiquadric', 'inverse_multiquadric',
             'spline', 'poisson_one', 'poisson_two', 'matern32', 'matern52']
    a = f32(np.linspace(0.0, 3.0, 61))
    for b, nm in enumerate(names):
        v, d = np.zeros_like(a), np.zeros_like(a)
        hc.hc_rbf(b, fp(a), fp(v), fp(d), len(a))
        ta = torch.tensor(a, dtype=torch.float64, requires_grad=True)
        tv = O._rbf_basis(nm)(ta)
        tv.sum().backward()
        assert_abs(v, npy(tv), 2e-6, nm)
        assert_abs(d, npy(ta.grad), 5e-6, nm + ' derivative')


@pytest.mark.parametrize('decoupled,wd', [(0, 0.0), (0, 1e-3), (1, 1e-2)])
def test_adam_matches_torch(hc, decoupled, wd):
    r = np.random.RandomState(3)
    p0 = f32(r.standard_normal(257))
    tp = torch.nn.Parameter(torch.tensor(p0))
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([tp], lr=1e-2, weight_decay=wd)
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for step in range(1, 4):
        g = f32(r.standard_normal(257))
        tp.grad = torch.tensor(g)
        opt.step()
        hc.hc_adam(fp(p), fp(g), fp(m), fp(v), len(p), ctypes.c_float(1e-2), ctypes.c_float(0.9),
                   ctypes.c_float(0.999), ctypes.c_float(1e-8), ctypes.c_float(wd), decoupled, step)
        assert_abs(p, npy(tp), 2e-6, 'adam step %d' % step)
