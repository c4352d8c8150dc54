"""Data-parallel sharding (nemo_cvpr2023_b200/dp.py) on CPU: two gloo ranks, each fitting half of the views with
the oracle, must reproduce the single-process gradients after one all-reduce of the shared parameters."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nemo_cvpr2023_b200 import assets, dp  # noqa: E402

V, FR = 4, 6


def _oracle(params, data, args):
    from oracle import nemo_oracle as O
    return O.FitOracle(args, O.to_torch(params), O.to_torch(assets.make_smpl_assets(0)),
                       O.to_torch(assets.make_vposer_weights(2)), O.to_torch(assets.make_gmm(3)), O.to_torch(data),
                       version=2).build_optimizers()


def _args():
    return assets.default_args(h_dim=32, phase_rbf_dim=8, instance_code_size=3, monotonic_network_n_nodes=8,
                               batch_size=V * FR)


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    torch.set_num_threads(1)
    args = _args()
    full_p, full_d = assets.make_fit_params(args, V, seed=7), assets.make_fit_data(V, FR, seed=1)
    views = dp.shard_views(V, world, rank)
    orc = _oracle(dp.shard_fit_params(full_p, views), dp.shard_fit_data(full_d, views), args)
    nv = len(views)
    vi, fi = torch.arange(nv).repeat_interleave(FR), torch.arange(FR).repeat(nv)
    out, _ = orc.losses(vi, fi)
    (out['total_loss'] / world).backward()
    shared = [t.grad for k, t in orc.p.items() if dp.is_shared(k) and t.grad is not None]
    dp.allreduce_shared_(shared)
    loss = out['total_loss'].detach().clone() / world
    dist.all_reduce(loss)
    res = {k: t.grad.numpy() for k, t in orc.p.items() if t.grad is not None}
    res['__loss__'] = loss.numpy()
    res['__views__'] = np.asarray(views)
    np.savez(out_path % rank, **res)
    dist.destroy_process_group()


def test_two_rank_shards_reproduce_single_process_gradients(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 500)
    out_path = str(tmp_path / 'rank%d.npz')
    mp.spawn(_worker, args=(world, port, out_path), nprocs=world, join=True)
    args = _args()
    full_p, full_d = assets.make_fit_params(args, V, seed=7), assets.make_fit_data(V, FR, seed=1)
    orc = _oracle(full_p, full_d, args)
    vi, fi = torch.arange(V).repeat_interleave(FR), torch.arange(FR).repeat(V)
    out, _ = orc.losses(vi, fi)
    out['total_loss'].backward()
    ranks = [dict(np.load(out_path % r)) for r in range(world)]
    for r in ranks:
        assert abs(float(r['__loss__']) - float(out['total_loss'])) <= 1e-5 * abs(float(out['total_loss']))
    for k, t in orc.p.items():
        if t.grad is None:
            continue
        g = t.grad.numpy()
        if dp.is_shared(k):
            for r in ranks:   # identical on every rank after the all-reduce, equal to the single-process gradient
                assert np.linalg.norm(r[k] - g) <= 1e-4 * max(np.linalg.norm(g), 1e-30), k
        elif k in ('learned_cameras', 'learned_instance_code'):
            for r in ranks:
                assert np.linalg.norm(r[k] - g[r['__views__']]) <= 1e-4 * max(np.linalg.norm(g[r['__views__']]), 1e-30), k
        elif k.startswith('phase_networks.'):
            v, nm = int(k.split('.')[1]), k.split('.')[2]
            for r in ranks:
                views = list(r['__views__'])
                if v in views:
                    lk = 'phase_networks.%d.%s' % (views.index(v), nm)
                    assert np.linalg.norm(r[lk] - g) <= 1e-4 * max(np.linalg.norm(g), 1e-30) + 1e-12, k


def test_shard_views_rejects_uneven_split():
    with pytest.raises(ValueError):
        dp.shard_views(5, 2, 0)
    assert dp.shard_views(8, 4, 3) == [6, 7]


def test_shard_plan_views_then_frames():
    """views first; with more ranks than views every view's frames are cut into equal windows (V < G)"""
    p = dp.shard_plan(8, 64, 4, 1)
    assert p == dict(views=[2, 3], frame_lo=0, frame_hi=64, reduce_all=False)
    plans = [dp.shard_plan(2, 64, 8, r) for r in range(8)]
    assert [q['views'] for q in plans] == [[0]] * 4 + [[1]] * 4
    assert [(q['frame_lo'], q['frame_hi']) for q in plans[:4]] == [(0, 16), (16, 32), (32, 48), (48, 64)]
    assert all(q['reduce_all'] for q in plans)
    cover = np.zeros((2, 64), int)
    for q in plans:
        cover[q['views'][0], q['frame_lo']:q['frame_hi']] += 1
    assert (cover == 1).all()
    with pytest.raises(ValueError):
        dp.shard_plan(3, 64, 8, 0)
    d = dp.shard_fit_data_frames(assets.make_fit_data(2, 64, seed=1), [1], 16, 32)
    assert d['points2d_gt_all'].shape == (1, 16, 25, 3) and d['gt_bbox_size'].shape == (1, 16) and d['IMG_D0'] == 1080


def _frame_split_worker(rank, world, port, out_path):
    """V = 1 view cut into two frame windows: every gradient (shared AND the view's own) is summed over the ranks and
    must equal the single-process gradient; the raw phase of a frame is its position in the WHOLE sequence."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    torch.set_num_threads(1)
    args = _args()
    full_p, full_d = assets.make_fit_params(args, 1, seed=7), assets.make_fit_data(1, FR * 2, seed=1)
    plan = dp.shard_plan(1, FR * 2, world, rank)
    assert plan['reduce_all']
    orc = _oracle(dp.shard_fit_params(full_p, plan['views']),
                  dp.shard_fit_data_frames(full_d, plan['views'], plan['frame_lo'], plan['frame_hi']), args)
    n = plan['frame_hi'] - plan['frame_lo']
    vi, fi = torch.zeros(n, dtype=torch.long), torch.arange(n)
    orc.raw_phase = lambda f: torch.linspace(0, 1, FR * 2)[f + plan['frame_lo']]
    out, _ = orc.losses(vi, fi)
    (out['total_loss'] / world).backward()
    grads = [t.grad for t in orc.p.values() if t.grad is not None]
    dp.allreduce_shared_(grads)
    np.savez(out_path % rank, **{k: t.grad.numpy() for k, t in orc.p.items() if t.grad is not None})
    dist.destroy_process_group()


def test_frame_split_reproduces_single_process_gradients(tmp_path):
    world = 2
    out_path = str(tmp_path / 'fs%d.npz')
    mp.spawn(_frame_split_worker, args=(world, 29100 + (os.getpid() % 300), out_path), nprocs=world, join=True)
    args = _args()
    orc = _oracle(assets.make_fit_params(args, 1, seed=7), assets.make_fit_data(1, FR * 2, seed=1), args)
    vi, fi = torch.zeros(FR * 2, dtype=torch.long), torch.arange(FR * 2)
    out, _ = orc.losses(vi, fi)
    out['total_loss'].backward()
    for r in range(world):
        res = dict(np.load(out_path % r))
        for k, t in orc.p.items():
            if t.grad is None or k == 'learned_betas':
                continue
            g = t.grad.numpy()
            assert np.linalg.norm(res[k] - g) <= 1e-4 * max(np.linalg.norm(g), 1e-30) + 1e-12, k
