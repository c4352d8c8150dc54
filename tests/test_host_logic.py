"""Host-side logic that needs no GPU: the optimiser-state mapping between torch.optim.Adam's per-parameter layout (what
the reference's `save` writes, nemo/neural_motion_model.py:257-265) and the engine's flat, padded, head-adjacent
buffers; the segment table of the fused Adam launch."""
import os

import numpy as np
import torch

from nemo_cvpr2023_b200 import fit_engine as fe
from nemo_cvpr2023_b200.nemo.neural_motion_model import FusedAdam
from oracle.make_golden import step_case
from tests.util import GOLD


def _motion_group(h, in_dim, D):
    """the motion optimiser's group as the engine builds it: a slice (offset != 0) of storage shared with other groups"""
    ldx = (in_dim + 3) // 4 * 4
    specs = fe.motion_specs(h, in_dim, ldx, D)
    n = fe.ParamGroup.flat_size(specs)
    flat = [torch.zeros(72 + n + 40) for _ in range(4)]
    g = fe.ParamGroup(specs, 'cpu', 1e-3, 1e-3, storage=tuple(t[72:72 + n] for t in flat))
    g.set_lagging(fe.LAGGING)
    g._flat = flat
    return g


def _torch_order(g):
    names = ['learned_motion.net.net.%d.%s' % (i, k) for i in (0, 2, 4) for k in ('weight', 'bias')]
    names += ['learned_motion.rot_out.weight', 'learned_motion.rot_out.bias', 'learned_motion.linear_out.weight',
              'learned_motion.linear_out.bias', 'phase_rbf.log_sigmas']
    return names, [g.views[n] for n in names]


def test_reference_optimizer_state_lands_in_the_flat_buffers():
    """opt_sd[1] of the reference-written checkpoint -> FusedAdam.load_state_dict -> every exp_avg / exp_avg_sq sits at
    the engine's offset of that tensor (row by row for the padded first-layer weight), pads stay zero, both step
    counters are set; state_dict() gives the torch layout back bit for bit."""
    args, _, _, _, _ = step_case(3)
    ck = torch.load(os.path.join(GOLD, 'ref_ckpt_v3.pt'), weights_only=False)
    sd = ck['opt_sd'][1]
    h, in_dim, D = args.h_dim, args.phase_rbf_dim + args.instance_code_size, args.phase_rbf_dim
    assert in_dim % 4 != 0   # the fixture exercises the padded row stride
    g = _motion_group(h, in_dim, D)
    names, params = _torch_order(g)
    opt = FusedAdam(params, g, 1e-3, 1e-3)
    opt.load_state_dict(sd)
    assert int(g.step_dev) == 5 and int(g.step_lag_dev) == 2
    used = torch.zeros_like(g.m, dtype=torch.bool)
    for i, n in enumerate(names):
        o, k = g.offsets[n]
        ref_m, ref_v = sd['state'][i]['exp_avg'], sd['state'][i]['exp_avg_sq']
        if n in g.padded:
            rows, ld = g.padded[n].shape
            mm, vv = g.m[o:o + k].view(rows, ld), g.v[o:o + k].view(rows, ld)
            assert torch.equal(mm[:, :ref_m.shape[1]], ref_m) and torch.equal(vv[:, :ref_m.shape[1]], ref_v), n
            assert float(mm[:, ref_m.shape[1]:].abs().sum()) == 0.0, n
            used[o:o + k].view(rows, ld)[:, :ref_m.shape[1]] = True
        else:
            assert torch.equal(g.m[o:o + k], ref_m.reshape(-1)) and torch.equal(g.v[o:o + k], ref_v.reshape(-1)), n
            used[o:o + k] = True
    assert float(g.m[~used].abs().sum()) == 0.0 and float(g.v[~used].abs().sum()) == 0.0
    for t in g._flat[2:]:   # nothing spilled into the neighbouring groups' slices
        assert float(t[:72].abs().sum()) == 0.0 and float(t[72 + g.m.numel():].abs().sum()) == 0.0
    back = opt.state_dict()
    assert sorted(back['state']) == sorted(sd['state'])
    for i in sd['state']:
        assert float(back['state'][i]['step']) == float(sd['state'][i]['step']), i
        assert torch.equal(back['state'][i]['exp_avg'], sd['state'][i]['exp_avg'])
        assert torch.equal(back['state'][i]['exp_avg_sq'], sd['state'][i]['exp_avg_sq'])
    assert back['param_groups'][0]['lr'] == sd['param_groups'][0]['lr']


def test_fresh_optimizer_has_no_state_and_warmup_leaves_lagging_tensors_without():
    g = _motion_group(48, 15, 12)
    names, params = _torch_order(g)
    opt = FusedAdam(params, g, 1e-3, 1e-3)
    assert opt.state_dict()['state'] == {}
    g.step_dev.fill_(3)   # three warm-up steps: linear_out.* (torch indices 8, 9) still have no state
    assert sorted(opt.state_dict()['state']) == [0, 1, 2, 3, 4, 5, 6, 7, 10]


def test_adam_segment_table():
    """ranges() covers the buffer exactly once in maximal runs; one segment per step counter owns the increment;
    skip_lag marks exactly the linear_out.* runs."""
    from nemo_cvpr2023_b200._lib import ADAM_OWNS_STEP, ADAM_SKIP
    g = _motion_group(1000, 105, 100)
    rs = g.ranges()
    assert [r[2] for r in rs] == [False, True, False, True, False]
    assert rs[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(rs, rs[1:])) and rs[-1][0] + rs[-1][1] == g.p.numel()
    assert all(r[0] % 4 == 0 for r in rs)
    assert rs[1][:2] == g.offsets['learned_motion.linear_out.weight'] and rs[3][0] == g.offsets['learned_motion.linear_out.bias'][0]
    segs = g.segments(skip_lag=True)
    assert [bool(s.flags & ADAM_SKIP) for s in segs] == [False, True, False, True, False]
    assert [bool(s.flags & ADAM_OWNS_STEP) for s in segs] == [True, True, False, False, False]
    assert sum(s.n for s in segs) == g.p.numel()
    plain = fe.ParamGroup([('learned_cameras', (8, 9))], 'cpu', 0.1)
    assert plain.ranges() == [(0, 72, False)] and plain.segments()[0].flags == ADAM_OWNS_STEP


def test_pending_losses_unpacks_the_packed_vector_once():
    """`step(..., sync=False)`: `PendingLosses.result()` waits for ITS copy's event, turns the packed [acc(8) | total]
    vector into the loss_dict of `step` (normalisers applied, requested keys only), and keeps the values after the pinned
    slot has been rewritten"""
    class Ev:
        waited = 0

        def synchronize(self):
            Ev.waited += 1

        def query(self):
            return False

    host = torch.tensor([6.0, 8.0, 0.25, 10.0, 3.0, 0.5, 0.0, 0.0, 42.0])
    norm = np.asarray([0.5, 0.25, 1.0, 0.1, 1.0, 1.0, 0.0, 0.0])
    keys = ['kp_loss', 'gmm_loss', 'vp_recon_loss', 'vp_kl_loss', 'total_loss']
    p = fe.PendingLosses(host, Ev(), norm, keys)
    assert not p.done()
    out = p.result()
    assert Ev.waited == 1 and p.done()
    assert list(out) == keys
    assert float(out['kp_loss']) == 3.0 and float(out['gmm_loss']) == 1.0 and float(out['vp_recon_loss']) == 3.0
    assert float(out['vp_kl_loss']) == 2.0 and float(out['total_loss']) == 42.0
    assert all(v.dtype == np.float32 for v in out.values())
    host.fill_(-1.0)                       # the slot is reused by a later step
    assert p.result() is out and float(p.result()['total_loss']) == 42.0 and Ev.waited == 1
    full = fe.PendingLosses(torch.arange(9.0), Ev(), norm).result()
    assert set(full) == {'kp_loss', 'vp_recon_loss', 'vp_kl_loss', 'gmm_loss', 'loss_3d', 'instance_loss', 'total_loss'}
