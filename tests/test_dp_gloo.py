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
