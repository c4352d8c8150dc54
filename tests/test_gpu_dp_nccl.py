"""N-rank GPU parity of the data-parallel fit ("the G-GPU result must equal the 1-GPU result up to fp32
summation order").  Two ranks over NCCL, one per GPU: an 8-view problem is cut with `dp.shard_plan`, every rank runs
`NemoV2.step` through `FitEngine.enable_data_parallel` (bucketed all-reduce on the communication stream), and the
gradients, the parameters after Adam and the reported losses are compared with the same step on ONE GPU.  Second case:
more ranks than views (V < G), where a view's frames are split and its camera / phase / instance gradients are summed
too.  Needs two GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp_nccl.py -m gpu`."""
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
from tests.util import assert_adam_param, assert_grad, assert_rel  # noqa: E402

pytestmark = pytest.mark.gpu
FR = 16


def _args(V):
    return assets.default_args(h_dim=256, phase_rbf_dim=20, instance_code_size=4, monotonic_network_n_nodes=24,
                               batch_size=V * FR, lr_human=1e-3)


def _problem(V):
    args = _args(V)
    params = assets.make_fit_params(args, V, seed=17)
    data = assets.make_fit_data(V, FR, seed=18)
    r = np.random.RandomState(19)
    data['points2d_gt_all'][..., 0] = 540 + 200 * r.standard_normal((V, FR, 25))
    data['points2d_gt_all'][..., 1] = 960 + 200 * r.standard_normal((V, FR, 25))
    return args, params, data


def _run(model, V, F):
    """one step (eager, bucketed all-reduce overlapped with the backward: gradients and parameters are compared after
    it), then two more -- the second call captures forward + backward and the optimisers into CUDA graphs around ONE
    eager all-reduce, the third replays them; their losses are returned too"""
    vi, fi = torch.arange(V).repeat_interleave(F), torch.arange(F).repeat(V)
    ld, _ = model.step(vi, fi)
    g = {k: v.detach().cpu().numpy().copy() for k, v in model.engine.named_grads().items()}
    p = {k: v.detach().cpu().numpy().copy() for k, v in model.engine.named_params().items()}
    out = {k: float(v) for k, v in ld.items()}
    for i in (2, 3):
        ld, _ = model.step(vi, fi)
        out['total_loss_step%d' % i] = float(ld['total_loss'])
    return out, g, p


def _worker(rank, world, port, V, out_path):
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = 'cuda:%d' % rank
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(dev))
    args, params, data = _problem(V)
    plan = dp.shard_plan(V, FR, world, rank)
    p = dp.shard_fit_params(params, plan['views'])
    d = dp.shard_fit_data_frames(data, plan['views'], plan['frame_lo'], plan['frame_hi'])
    m = nmm.NemoV2(args, SyntheticMultiViewSequence(data=d), dev, init_params=p)
    Fl = plan['frame_hi'] - plan['frame_lo']
    if Fl != FR:
        m.engine.set_frame_window(plan['frame_lo'], FR)
    m.engine.enable_data_parallel(world, reduce_all=plan['reduce_all'])
    ld, g, pp = _run(m, len(plan['views']), Fl)
    np.savez(out_path % rank, views=np.asarray(plan['views']), **{'loss.' + k: v for k, v in ld.items()},
             **{'grad.' + k: v for k, v in g.items()}, **{'param.' + k: v for k, v in pp.items()})
    dist.destroy_process_group()


def _local_name(k, views):
    """name of a full-problem tensor on a rank that holds `views` (None if the rank does not hold it)"""
    if k.startswith('phase_networks.'):
        v = int(k.split('.')[1])
        return 'phase_networks.%d.%s' % (views.index(v), k.split('.')[2]) if v in views else None
    return k


@pytest.mark.timeout(240)
@pytest.mark.parametrize('V,symm', [(8, 0), (1, 0), (8, 2)])
def test_two_ranks_equal_one_gpu(tmp_path, monkeypatch, V, symm):
    """symm = 2: the shared gradient is summed by this library's peer-memory kernels inside the step graph
    (csrc/allreduce_p2p.cu, NEMO_B200_DP_SYMM=2) instead of NCCL"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    monkeypatch.setenv('NEMO_B200_DP_SYMM', str(symm))    # read by the spawned ranks when they import the engine
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence
    world = 2
    args, params, data = _problem(V)
    full = nmm.NemoV2(args, SyntheticMultiViewSequence(data=data), 'cuda:0', init_params=params)
    ld, g, p = _run(full, V, FR)
    del full
    torch.cuda.empty_cache()
    out_path = str(tmp_path / 'rank%d.npz')
    ctx = mp.spawn(_worker, args=(world, 29600 + (os.getpid() % 300), V, out_path), nprocs=world, join=False)
    import time
    t0 = time.time()
    while not ctx.join(timeout=5):
        if time.time() - t0 > 180:   # a hung collective must not hold the GPU box: kill the ranks and fail
            for p in ctx.processes:
                p.kill()
            pytest.fail('data-parallel ranks did not finish within 180 s')
    for r in range(world):
        res = dict(np.load(out_path % r))
        views = list(res['views'])
        for k in ld:   # every rank reports the GLOBAL losses
            # steps 2 / 3 (captured / replayed) run on parameters that one Adam step has moved by ~lr in every entry
            assert_rel(res['loss.' + k], ld[k], 2e-3 if 'step' in k else 1e-5, 'rank %d %s' % (r, k))
        n = 0
        for k in g:
            if k == 'learned_betas':
                continue
            lk = _local_name(k, views)
            if lk is None:
                continue
            ref_g, ref_p = g[k], p[k]
            if k in ('learned_cameras', 'learned_instance_code'):
                ref_g, ref_p = ref_g[views], ref_p[views]
            # shared tensors: summed over the ranks; per-view tensors: already scaled for the global loss
            assert_grad(res['grad.' + lk], ref_g, what='rank %d grad %s' % (r, k))
            assert_adam_param(res['param.' + lk], ref_p, 'rank %d param %s' % (r, k),
                              lr=args.lr_camera if 'camera' in k else 1e-3)
            n += 1
        assert n >= 15
