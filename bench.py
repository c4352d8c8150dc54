#!/usr/bin/env python
"""Benchmark of the NeMo per-iteration fitting hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 0..4] [--impl reference]

A "step" is one `NemoV*.step` (motion MLP -> SMPL LBS -> reprojection loss, + VPoser and GMM priors, forward +
backward + Adam, exactly what the reference's `step` does).  The default workload (the headline, `--config 1`) is a
batch of 4096 (view, frame) samples per GPU -- the batch of BASELINE.json configs[1] through the full step of
configs[0], weak scaling.  `--config 0 / 2 / 3 / 4` run the other BASELINE configurations (strong scaling: the views
are cut over the ranks).  Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from nemo_cvpr2023_b200 import assets  # noqa: E402

METRIC = 'fit frames/s fwd+bwd (MLP->SMPL->reproj loss)'
STREAM_BYTES = 6890 * 3 * 4           # one per-vertex fp32 stream of one frame

# BASELINE.json configs.  V views x F frames per action; `weak`: V is per GPU, else the views are cut over the ranks.
CONFIGS = {
    0: dict(tag='configs[0]', version=2, actions=1, V=8, F=64, h=1000, weak=False,
            what='NeMo baseball-pitch fit: 1 action x 8 video instances x 64 frames'),
    1: dict(tag='configs[1] batch through the configs[0] step', version=2, actions=1, V=8, F=512, h=1000, weak=True,
            what='8 views x 512 frames = 4096 samples per GPU per step'),
    2: dict(tag='configs[2]', version=2, actions=8, V=8, F=128, h=1000, weak=False,
            what='NeMo-MoCap scale fit: 8 actions (8 parameter sets) x 8 instances x 128 frames, instances cut over the '
                 'ranks, ONE all-reduce of the 8 x 9 MB shared gradients'),
    3: dict(tag='configs[3]', version=4, actions=1, V=64, F=512, h=1024, weak=False,
            over=dict(weight_3d_loss=1000.0),   # the reference's V3 / V4 step only runs with a 3D-loss weight (:3892)
            what='large-batch fit: 64 instances x 512 frames = 32768 samples, motion MLP width 1024, per-instance '
                 'trainable cameras (NemoV4)'),
    4: dict(tag='configs[4]', version=3, actions=1, V=16, F=256, h=1000, weak=False,
            over=dict(weight_3d_loss=1000.0, weight_instance_loss=0.1),
            what='custom-video path: NemoV3, 16 instances x 256 frames full batch, 2D OpenPose + 3D pose losses, '
                 'per-instance codes'),
}


def bench_args(**over):
    d = dict(h_dim=1000, phase_rbf_dim=100, instance_code_size=5, monotonic_network_n_nodes=200,
             weight_vp_loss=10.0, weight_vp_z_loss=1.0, weight_gmm_loss=1.0, loss='mse_robust')
    d.update(over)
    return assets.default_args(**d)


def config_args(c, n):
    return bench_args(batch_size=n, h_dim=c['h'], **c.get('over', {}))


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms by one long-running `nvidia-smi -lms`, started
    before warm-up; only samples that fall inside a timed `window()` count (the GPU is under load there)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows, self.windows, self.index, self.proc = [], [], index, None

    def _run(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.strip().split(',')]))

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.t.join(timeout=5)
            self.proc = None

    @contextlib.contextmanager
    def window(self):
        t0 = time.perf_counter()
        try:
            yield self
        finally:
            self.windows.append((t0, time.perf_counter()))

    def summary(self):
        rows = [r for t, r in self.rows if len(r) >= 6 and r[0].isdigit()
                and any(a <= t <= b for a, b in self.windows)]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith('active') for r in rows)]
        return {'sm_mhz': float(np.median([float(r[0]) for r in rows])), 'sm_max_mhz': float(rows[0][1]),
                'reasons': reasons, 'samples': len(rows)}


def workload_config(cfg_id, gpus):
    c = CONFIGS[cfg_id]
    per_gpu = c['actions'] * c['V'] * c['F'] if c['weak'] else c['actions'] * c['V'] * c['F'] // gpus
    return {'workload': 'Nemo%s.step fwd+bwd+Adam (motion MLP h=%d, RBF 100, code 5 -> SMPL LBS 6890 verts / 24 joints / '
                        '10 betas -> projection -> mse_robust 2D loss + VPoser v2v/KL + GMM prior): %s (%s)'
                        % ('V%d' % c['version'], c['h'], c['what'], c['tag']),
            'baseline_config': cfg_id, 'samples_per_gpu': per_gpu, 'global_batch': per_gpu * gpus,
            'l2_policy': 'per-step working set (340 MB pose-offset streams + gradients at 4096 samples) exceeds the '
                         '126 MB L2' if per_gpu >= 2048 else 'working set of this small configuration fits the L2; every '
                         'step rewrites all of it',
            'parallelism': 'dp%d over video instances (views); NCCL all-reduce of the shared MLP gradients, bucketed and '
                           'overlapped with the backward' % gpus}


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_problem(c):
    """The step of configuration `c` on ONE GPU's share at --gpus 1 (weak configs: one GPU's batch)."""
    V, F = c['V'], c['F']
    V = max(1, min(V, 4096 // F))        # bounded sample: at most 4096 samples per CPU step (configs[3]: 8 of its 64 views)
    args = config_args(c, V * F)
    params = assets.make_fit_params(args, V, seed=4)
    data = assets.make_fit_data(V, F, seed=1)
    vi = torch.arange(V).repeat_interleave(F)
    fi = torch.arange(F).repeat(V)
    return args, params, data, vi, fi


def make_cpu_step(c):
    """-> (callable running one step, kind).  kind "reference": the UNMODIFIED reference's `NemoV*.step` from
    oracle/_ref (or /root/reference where mounted); kind "port": the oracle restatement, when no reference copy is
    present."""
    from oracle import ref_harness as rh
    args, params, data, vi, fi = cpu_problem(c)
    sm, vp, gmm = assets.make_smpl_assets(0), assets.make_vposer_weights(2), assets.make_gmm(3)
    if rh.available():
        try:
            with contextlib.redirect_stdout(sys.stderr):
                mods = rh.setup(sm)
                m = rh.build_fit_model(mods, args, params, sm, vp, gmm, data, version=c['version'])

            def step():
                with contextlib.redirect_stdout(sys.stderr):
                    return m.step(vi, fi)[0]
            return step, 'reference'
        except Exception as e:  # a broken copy must not take the bench down: say so and time the port
            sys.stderr.write('reference import failed (%r); timing the oracle port\n' % (e,))
    from oracle import nemo_oracle as O
    orc = O.FitOracle(args, O.to_torch(params), O.to_torch(sm), O.to_torch(vp), O.to_torch(gmm), O.to_torch(data),
                      version=c['version']).build_optimizers()
    return (lambda: orc.step(vi, fi)[0]), 'port'


def time_cpu(c, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind = make_cpu_step(c)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    med = float(np.median(ts))
    n = max(1, min(c['V'], 4096 // c['F'])) * c['F']
    return n / med, med, torch.get_num_threads(), kind, n


def cpu_sample_text(c, kind, n, warmup, steps, med):
    who = ("the unmodified reference's Nemo%s.step (oracle/_ref, PyTorch CPU fp32)" % ('V%d' % c['version'])
           if kind == 'reference' else 'oracle port of the reference step (PyTorch CPU fp32)')
    return ('%s, all host threads, on %d views x %d frames = %d samples per step (one action model of the GPU arm\'s '
            'configuration, h=%d, all priors + Adam); %d warm-up + %d timed steps, median %.0f ms'
            % (who, n // c['F'], c['F'], n, c['h'], warmup, steps, med * 1e3))


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    c = CONFIGS[a.config]
    steps, warmup = max(1, min(a.steps, 5)), max(1, min(a.warmup, 1))
    fps, med, cores, kind, n = time_cpu(c, steps, warmup)
    line = {'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': a.gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': med * 1e3, 'higher_is_better': True,
            'scaling': 'weak' if c['weak'] else 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(a.config, a.gpus),
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': kind,
                             'sample': cpu_sample_text(c, kind, n, warmup, steps, med)},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def _time_events(fn, iters):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def k7_roofline(dev, N, iters=10):
    """Skinning kernels timed alone with CUDA events on the current stream.
    pair: forward (read pose offsets, write vertices) + backward (read pose offsets, read dV, write d pose offsets)
          = 5 per-vertex streams per frame -- the figure SURVEY.md 8(d) defines.
    in_step: the two variants the fit step launches -- fused L1 of two meshes (read 2 pose-offset streams, write sign
          bits) + backward from the sign bits (read pose offsets, write d pose offsets) = 4 streams per frame."""
    from nemo_cvpr2023_b200 import functional as Fn
    from nemo_cvpr2023_b200._lib import lib
    C = Fn.SMPLConstants(assets.make_smpl_assets(0), dev)
    g = torch.Generator(device=dev).manual_seed(0)
    A = torch.randn(N, 24, 12, device=dev, generator=g)
    A2 = A + 0.01 * torch.randn(N, 24, 12, device=dev, generator=g)
    PO = 0.01 * torch.randn(N, 3 * Fn.VP, device=dev, generator=g)
    PO2 = 0.01 * torch.randn(N, 3 * Fn.VP, device=dev, generator=g)
    dV = torch.randn(N, Fn.NV, 3, device=dev, generator=g)
    Vout = torch.empty(N, Fn.NV, 3, device=dev)
    dPO, dAs = torch.empty_like(PO), torch.empty(N, 24, 12, device=dev)
    signs = torch.zeros(N, 3, 2, Fn.VP // 32, dtype=torch.int32, device=dev)
    l1 = torch.zeros(1, device=dev)
    out = {}
    out['fwd'] = _time_events(lambda: lib.skin_fwd(A, PO, C.vt, C.Wt, Vout, N), iters)
    out['bwd'] = _time_events(lambda: lib.skin_bwd(A, PO, C.vt, C.Wt, dV, None, 0.0, dPO, dAs, N), iters)
    out['l1_fwd'] = _time_events(lambda: lib.skin_l1_fwd(A, PO, A2, PO2, C.vt, C.Wt, signs, l1, N), iters)
    out['bwd_signs'] = _time_events(lambda: lib.skin_bwd(A, PO, C.vt, C.Wt, None, signs, 1e-6, dPO, dAs, N), iters)
    return out


def k7_traffic(n):
    """dram__bytes_read + dram__bytes_write of the K7 launch pair from the committed ncu --set full capture (None
    when the capture was taken at another batch size); the capture's file and commit are named in the JSON."""
    p = os.path.join(ROOT, 'profiles', 'k7_traffic.json')
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    if d.get('n') != n:
        return None, None
    return d['fwd_dram_bytes'] + d['bwd_dram_bytes'], d.get('source', 'profiles/k7_traffic.json')


def count_kernels(fn):
    """Every GPU kernel (this library's AND ATen's fills / copies / NCCL) one call of `fn` launches, from CUPTI."""
    try:
        from torch.profiler import ProfilerActivity, profile
        fn()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        names = [e.name for e in prof.events() if getattr(e, 'device_type', None) is not None
                 and str(e.device_type).endswith('CUDA') and not e.name.lower().startswith(('memcpy', 'memset'))]
        mem = [e.name for e in prof.events() if str(getattr(e, 'device_type', '')).endswith('CUDA')
               and e.name.lower().startswith(('memcpy', 'memset'))]
        lib_marks = ('at::', 'nccl', 'cub::', 'thrust::', 'c10::')
        repo = [n for n in names if not any(mk in n for mk in lib_marks)]
        return {'kernels': len(names), 'repo': len(repo), 'other': len(names) - len(repo), 'memops': len(mem)}
    except Exception as e:  # CUPTI unavailable: the library's own count stands
        return {'error': repr(e)[:120]}


def parity_check(dev):
    """The GPU step and the CPU oracle on the SAME batch (configs[0]: 8 x 64 = 512 samples, h = 1000): worst relative
    error over the loss terms and worst per-tensor relative L2 error over every parameter gradient."""
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence
    from oracle import nemo_oracle as O
    c = CONFIGS[0]
    args, params, data, vi, fi = cpu_problem(c)
    orc = O.FitOracle(args, O.to_torch(params), O.to_torch(assets.make_smpl_assets(0)),
                      O.to_torch(assets.make_vposer_weights(2)), O.to_torch(assets.make_gmm(3)), O.to_torch(data),
                      version=2).build_optimizers()
    ref, _ = orc.step(vi, fi)
    m = nmm.NemoV2(args, SyntheticMultiViewSequence(data=data), dev, init_params=params)
    ld, _ = m.step(vi, fi)
    loss_err = max(abs(float(ld[k]) - ref[k]) / max(abs(ref[k]), 1e-30) for k in ld)
    grads = m.engine.named_grads()
    gerr, worst = 0.0, ''
    for k, p in orc.p.items():
        if p.grad is None or k == 'learned_betas' or float(p.grad.abs().max()) == 0.0:
            continue
        g, r = grads[k].detach().cpu().double().reshape(-1), p.grad.double().reshape(-1)
        e = float((g - r).norm() / r.norm())
        if e > gerr:
            gerr, worst = e, k
    return {'against': 'CPU oracle (oracle/nemo_oracle.py) on the same batch: configs[0], 512 samples, h=1000',
            'max_rel_loss_err': loss_err, 'max_rel_l2_grad_err': gerr, 'worst_grad': worst,
            'tolerance': {'loss': 1e-5, 'grad': 1e-4}, 'ok': bool(loss_err <= 1e-5 and gerr <= 1e-4)}


def run_gpu(a):
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    if world > 1:
        import torch.distributed as dist
        # stdout carries ONE JSON line and nothing else: NCCL prints its version banner there when NCCL_DEBUG=VERSION
        # (set in this image), so file descriptor 1 points at stderr while the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device(dev))
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    try:
        _run_gpu(a, rank, local, world, dev)
    finally:
        if world > 1:
            # NCCL_DEBUG=INFO: the communicator's teardown messages ("Destroy COMPLETE", every rank) must not follow the
            # JSON line on stdout
            sys.stdout.flush()
            os.dup2(2, 1)
            torch.distributed.destroy_process_group()


def _run_gpu(a, rank, local, world, dev):
    from nemo_cvpr2023_b200 import dp
    from nemo_cvpr2023_b200._lib import lib
    from nemo_cvpr2023_b200.multi_action import MultiActionFit
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence

    c = CONFIGS[a.config]
    cls = {2: nmm.NemoV2, 3: nmm.NemoV3, 4: nmm.NemoV4}[c['version']]
    F_total = c['F']
    if c['weak']:
        plan = dict(views=list(range(c['V'])), frame_lo=0, frame_hi=F_total, reduce_all=False)
        V_glob = c['V']
    else:
        plan = dp.shard_plan(c['V'], F_total, world, rank)
        V_glob = c['V']
    Vl, Fl = len(plan['views']), plan['frame_hi'] - plan['frame_lo']
    N = Vl * Fl                                   # samples per action model per step on this rank
    args = config_args(c, N)
    models = []
    for act in range(c['actions']):
        # shared (data-parallel) parameters start identical on every rank; per-view ones belong to the view
        full_p = assets.make_fit_params(args, V_glob, seed=4 + 100 * act + (rank if c['weak'] else 0))
        shared = assets.make_fit_params(args, V_glob, seed=4 + 100 * act)
        for k in shared:
            if dp.is_shared(k):
                full_p[k] = shared[k]
        full_d = assets.make_fit_data(V_glob, F_total, seed=1 + 100 * act + (rank if c['weak'] else 0))
        p = dp.shard_fit_params(full_p, plan['views'])
        d = dp.shard_fit_data_frames(full_d, plan['views'], plan['frame_lo'], plan['frame_hi'])
        m = cls(args, SyntheticMultiViewSequence(data=d), dev, init_params=p)
        if Fl != F_total:
            m.engine.set_frame_window(plan['frame_lo'], F_total)
        models.append(m)
    multi = MultiActionFit(models, dev) if c['actions'] > 1 else None
    if world > 1:
        if multi:
            multi.enable_data_parallel(world, reduce_all=plan['reduce_all'])
        else:
            models[0].engine.enable_data_parallel(world, reduce_all=plan['reduce_all'])
    vi_h = torch.arange(Vl).repeat_interleave(Fl).pin_memory()
    fi_h = torch.arange(Fl).repeat(Vl).pin_memory()
    vi_d, fi_d = vi_h.to(dev), fi_h.to(dev)
    A = c['actions']

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, flush=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        if flush:
            flush()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t)

    model = models[0]
    eng = model.engine
    train = c['version'] >= 3

    if multi:
        def dev_step():
            multi.train_step([vi_d] * A, [fi_d] * A)

        def e2e_step():
            multi.step([vi_h] * A, [fi_h] * A)

        e2e_flush = None
    else:
        def dev_step():      # inputs resident in HBM, no host read-back; one CUDA-graph launch on one GPU
            eng.training = train
            eng.step_graph(vi_d, fi_d, mode='step')

        def e2e_blocking():  # the public call: pinned host indices in, loss dict out, the host waits for it
            model.step(vi_h, fi_h)

        pending = []

        def e2e_step():
            # the same call with the read-back pipelined by one step, like a fit loop that logs: the losses of every step
            # are copied to pinned host memory and READ by the host, those of step i while step i + 1 runs
            out, _ = model.step(vi_h, fi_h, sync=False)
            if pending:
                pending.pop().result()
            pending.append(out)

        def e2e_flush():     # the last step's losses are read inside the timed region too
            while pending:
                pending.pop().result()

    cs = ClockSampler(local).start()
    warm = max(3, a.warmup)
    for _ in range(warm):
        dev_step()
    with cs.window():
        l0 = lib.launches
        t_dev = timed(dev_step, a.steps)
        launches = (lib.launches - l0) // a.steps
    t_e2e_blocking = None
    if not multi:
        for _ in range(2):
            e2e_blocking()
        t_e2e_blocking = timed(e2e_blocking, a.steps)
    for _ in range(2):
        e2e_step()
    if e2e_flush:
        e2e_flush()
    with cs.window():
        t_e2e = timed(e2e_step, a.steps, e2e_flush)
    kernels = count_kernels(dev_step)     # every rank: the step holds a collective
    if rank != 0:
        cs.stop()
        return
    total = A * N * world
    value = total * a.steps / t_dev
    e2e_val = total * a.steps / t_e2e
    peak, peak_src = peaks()
    nK7 = 4096
    with cs.window():
        tk = k7_roofline(dev, nK7)
    cs.stop()
    clocks = cs.summary()
    alg = 5 * STREAM_BYTES * nK7
    achieved = alg / (tk['fwd'] + tk['bwd']) / 1e9
    alg_in = 4 * STREAM_BYTES * nK7
    in_step = alg_in / (tk['l1_fwd'] + tk['bwd_signs']) / 1e9
    traffic, traffic_src = k7_traffic(nK7)
    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': a.steps, 'warmup': warm,
        'ms_per_step': t_dev / a.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak' if c['weak'] else 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(a.config, world),
        'clocks': clocks,
        'e2e': {'value': e2e_val, 'unit': 'frames/s', 'h2d_bytes_per_step': int(A * 2 * N * 8),
                'd2h_bytes_per_step': A * 9 * 4, 'ms_per_step': t_e2e / a.steps * 1e3,
                'readback': ('blocking' if multi else 'every step\'s packed loss vector is copied to pinned host memory '
                             'and read by the host one step later (model.step(..., sync=False) -> PendingLosses.result())'),
                'blocking_ms_per_step': (t_e2e_blocking / a.steps * 1e3 if t_e2e_blocking else None)},
        'gpu_launches': int(kernels.get('kernels', launches)),
        'gpu_launches_detail': dict(kernels, repo_c_abi_count=int(launches)),
        'roofline': {'bound': 'hbm', 'kernel': 'k_skin_fwd2<0> + k_skin_bwd2<dV> (K7, vertices written, dV read: 5 streams/frame)',
                     'achieved': achieved, 'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                     'frac': achieved / peak, 'traffic': traffic, 'traffic_source': traffic_src,
                     'algorithmic_bytes_per_launch_pair': alg, 'frames_per_launch': nK7,
                     'fwd_ms': tk['fwd'] * 1e3, 'bwd_ms': tk['bwd'] * 1e3,
                     'fwd_frac': 2 * STREAM_BYTES * nK7 / tk['fwd'] / 1e9 / peak,
                     'bwd_frac': 3 * STREAM_BYTES * nK7 / tk['bwd'] / 1e9 / peak,
                     'in_step': {'kernel': 'k_skin_fwd2<1> (fused L1 of two meshes) + k_skin_bwd2<signs>: the variants '
                                           'the timed step launches, 4 streams/frame',
                                 'achieved': in_step, 'frac': in_step / peak, 'algorithmic_bytes_per_launch_pair': alg_in,
                                 'l1_fwd_ms': tk['l1_fwd'] * 1e3, 'bwd_signs_ms': tk['bwd_signs'] * 1e3}},
    }
    if world == 1:
        line['parity'] = parity_check(dev)
        cpu_fps, cpu_med, cores, kind, n_cpu = time_cpu(c, 3, 1)
        line['cpu_baseline'] = {'value': cpu_fps, 'unit': 'frames/s', 'cores': cores, 'kind': kind,
                                'sample': cpu_sample_text(c, kind, n_cpu, 1, 3, cpu_med)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--config', type=int, default=1, choices=sorted(CONFIGS))
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_gpu(a)


if __name__ == '__main__':
    main()
