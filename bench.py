#!/usr/bin/env python
"""Benchmark of the NeMo per-iteration fitting hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one NemoV2.step (motion MLP -> SMPL LBS -> reprojection loss, + VPoser and GMM priors,
forward + backward + Adam, exactly what the reference's `step` does) on a batch of 4096 (view, frame)
samples per GPU -- the batch shape of BASELINE.json configs[1] through the full step of configs[0].
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
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
V_PER_GPU, F_FRAMES = 8, 512          # 4096 samples per GPU per step
CPU_V, CPU_F = 8, 64                  # bounded CPU sample: configs[0], 512 samples per step
STREAM_BYTES = 6890 * 3 * 4           # one per-vertex fp32 stream of one frame


def bench_args(**over):
    d = dict(h_dim=1000, phase_rbf_dim=100, instance_code_size=5, monotonic_network_n_nodes=200,
             weight_vp_loss=10.0, weight_vp_z_loss=1.0, weight_gmm_loss=1.0, loss='mse_robust')
    d.update(over)
    return assets.default_args(**d)


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


# ------------------------------------------------------------------------------------------ CPU arms
def make_oracle(V, F, seed=0):
    from oracle import nemo_oracle as O
    args = bench_args(batch_size=V * F)
    params = assets.make_fit_params(args, V, seed=4)
    data = assets.make_fit_data(V, F, seed=1)
    orc = O.FitOracle(args, O.to_torch(params), O.to_torch(assets.make_smpl_assets(0)),
                      O.to_torch(assets.make_vposer_weights(2)), O.to_torch(assets.make_gmm(3)), O.to_torch(data),
                      version=2).build_optimizers()
    vi = torch.arange(V).repeat_interleave(F)
    fi = torch.arange(F).repeat(V)
    return orc, vi, fi


def time_cpu(steps, warmup):
    """The reference's algorithm (oracle port, PyTorch CPU, all host threads) on the bounded sample."""
    torch.set_num_threads(os.cpu_count() or 1)
    orc, vi, fi = make_oracle(CPU_V, CPU_F)
    for _ in range(warmup):
        orc.step(vi, fi)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.step(vi, fi)
        ts.append(time.perf_counter() - t0)
    med = float(np.median(ts))
    return CPU_V * CPU_F / med, med, torch.get_num_threads()


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warmup = max(1, min(a.steps, 5)), max(1, min(a.warmup, 2))
    fps, med, cores = time_cpu(steps, warmup)
    sample = ('oracle port of the reference step (PyTorch CPU fp32), NemoV2.step on %d views x %d frames = %d '
              'samples per step, h=1000, all priors + Adam; %d warm-up + %d timed steps, median'
              % (CPU_V, CPU_F, CPU_V * CPU_F, warmup, steps))
    line = {'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': a.gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': med * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(a.gpus),
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def workload_config(gpus):
    return {'workload': 'NemoV2.step fwd+bwd+Adam (motion MLP h=1000, RBF 100, code 5 -> SMPL LBS 6890 verts / 24 joints '
                        '/ 10 betas -> projection -> mse_robust 2D loss + VPoser v2v/KL + GMM prior), %d views x %d frames '
                        '= %d samples per GPU per step (configs[1] batch through the configs[0] step)'
                        % (V_PER_GPU, F_FRAMES, V_PER_GPU * F_FRAMES),
            'samples_per_gpu': V_PER_GPU * F_FRAMES, 'global_batch': V_PER_GPU * F_FRAMES * gpus,
            'l2_policy': 'per-step working set (3 x 340 MB pose-offset streams + gradients) exceeds the 126 MB L2',
            'parallelism': 'dp%d over video instances (views); NCCL allreduce of shared MLP gradients' % gpus}


# ------------------------------------------------------------------------------------------ GPU arm
def k7_roofline(dev, iters=10):
    """Skinning kernel pair timed alone with CUDA events: forward (read pose offsets, write vertices) +
    backward (read pose offsets, read dV, write d pose offsets) = 5 per-vertex streams per frame."""
    from nemo_cvpr2023_b200 import functional as Fn
    from nemo_cvpr2023_b200._lib import lib
    N = V_PER_GPU * F_FRAMES
    C = Fn.SMPLConstants(assets.make_smpl_assets(0), dev)
    g = torch.Generator(device=dev).manual_seed(0)
    A = torch.randn(N, 24, 12, device=dev, generator=g)
    PO = 0.01 * torch.randn(N, 3 * Fn.VP, device=dev, generator=g)
    dV = torch.randn(N, Fn.NV, 3, device=dev, generator=g)
    Vout = torch.empty(N, Fn.NV, 3, device=dev)
    dPO, dAs = torch.empty_like(PO), torch.empty(N, 24, 12, device=dev)

    def fwd():
        lib.skin_fwd(A, PO, C.vt, C.Wt, Vout, N)

    def bwd():
        lib.skin_bwd(A, PO, C.vt, C.Wt, dV, None, 0.0, dPO, dAs, N)

    out = {}
    for name, fn in (('fwd', fwd), ('bwd', bwd)):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / iters * 1e-3
    return N, out


def k7_traffic(n):
    """dram__bytes_read + dram__bytes_write of the K7 launch pair from the committed ncu --set full capture (None
    when the capture was taken at another batch size)."""
    p = os.path.join(ROOT, 'profiles', 'k7_traffic.json')
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d['fwd_dram_bytes'] + d['bwd_dram_bytes'] if d.get('n') == n else None


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
    from nemo_cvpr2023_b200._lib import lib
    from nemo_cvpr2023_b200.nemo import neural_motion_model as nmm
    from nemo_cvpr2023_b200.nemo.multi_view_sequence import SyntheticMultiViewSequence

    N = V_PER_GPU * F_FRAMES
    args = bench_args(batch_size=N)
    params = assets.make_fit_params(args, V_PER_GPU, seed=4 + rank)
    # shared (data-parallel) parameters start identical on every rank
    shared = assets.make_fit_params(args, V_PER_GPU, seed=4)
    for k in shared:
        if k.startswith('learned_motion.') or k.startswith('phase_rbf.'):
            params[k] = shared[k]
    model = nmm.NemoV2(args, SyntheticMultiViewSequence(V_PER_GPU, F_FRAMES, seed=1 + rank), dev, init_params=params)
    if world > 1:
        model.engine.enable_data_parallel(world)
    vi_h = torch.arange(V_PER_GPU).repeat_interleave(F_FRAMES).pin_memory()
    fi_h = torch.arange(F_FRAMES).repeat(V_PER_GPU).pin_memory()
    vi_d, fi_d = vi_h.to(dev), fi_h.to(dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t)

    eng = model.engine

    def dev_step():      # inputs resident in HBM, no host read-back
        eng.training = False
        eng.forward_backward(vi_d, fi_d, update=True, mode='step')
        eng.optimizer_step()

    def e2e_step():      # the public call: pinned host indices in, loss dict out
        model.step(vi_h, fi_h)

    cs = ClockSampler(local).start()
    for _ in range(max(3, a.warmup)):
        dev_step()
    with cs.window():
        l0 = lib.launches
        t_dev = timed(dev_step, a.steps)
        launches = (lib.launches - l0) // a.steps
    for _ in range(2):
        e2e_step()
    with cs.window():
        t_e2e = timed(e2e_step, a.steps)
    if rank != 0:
        cs.stop()
        return
    total = N * world
    value = total * a.steps / t_dev
    e2e_val = total * a.steps / t_e2e
    peak, peak_src = peaks()
    with cs.window():
        nK7, tk = k7_roofline(dev)
    cs.stop()
    clocks = cs.summary()
    alg = 5 * STREAM_BYTES * nK7
    achieved = alg / (tk['fwd'] + tk['bwd']) / 1e9
    cpu_fps, cpu_med, cores = time_cpu(3, 1)
    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': a.steps, 'warmup': max(3, a.warmup),
        'ms_per_step': t_dev / a.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(world), 'clocks': clocks,
        'e2e': {'value': e2e_val, 'unit': 'frames/s', 'h2d_bytes_per_step': int(2 * N * 8),
                'd2h_bytes_per_step': 9 * 4, 'ms_per_step': t_e2e / a.steps * 1e3},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'kernel': 'k_skin_tc_fwd<0> + k_skin_tc_bwd<dV> (K7, vertices written, dV read: 5 streams/frame)',
                     'achieved': achieved, 'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                     'frac': achieved / peak, 'traffic': k7_traffic(nK7),
                     'algorithmic_bytes_per_launch_pair': alg, 'fwd_ms': tk['fwd'] * 1e3, 'bwd_ms': tk['bwd'] * 1e3,
                     'fwd_frac': 2 * STREAM_BYTES * nK7 / tk['fwd'] / 1e9 / peak,
                     'bwd_frac': 3 * STREAM_BYTES * nK7 / tk['bwd'] / 1e9 / peak},
        'cpu_baseline': {'value': cpu_fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                         'sample': 'oracle port of NemoV2.step (PyTorch CPU fp32, all host threads) on %d views x %d '
                                   'frames = %d samples per step, 1 warm-up + 3 timed steps, median %.0f ms'
                                   % (CPU_V, CPU_F, CPU_V * CPU_F, cpu_med * 1e3)},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_gpu(a)


if __name__ == '__main__':
    main()
