"""Several action models fitted side by side on one GPU (BASELINE.json configs[2]: 8 actions x 8 instances x 128 frames).

The reference fits ONE action per process and loops over the actions in a shell script
(`run_scripts_examples/nemomocap-example.sh:1-2,23-47`): the action models share nothing but the frozen SMPL / VPoser /
GMM assets.  Here every action keeps its own `NemoV*` model (own parameter set, own optimisers, the reference's
state_dict per action) and one iteration steps all of them: each action's kernels are enqueued on its own CUDA stream,
so the small per-action launches (1024 samples: 72 GEMM tiles for 148 SMs) fill the GPU together, and the optimiser
updates of all actions go out as one fused Adam launch per action model.  Data parallel: every rank holds all action
models and a slice of each action's views; the shared MLP gradients of ALL actions are summed in ONE all-reduce
(A x 9 MB) before the Adam launches.
"""
from __future__ import annotations

import numpy as np
import torch

from .nemo.neural_motion_model import _StepInfo


class MultiActionFit:
    def __init__(self, models, device=None):
        self.models = list(models)
        self.dev = torch.device(device if device is not None else self.models[0].engine.dev)
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in self.models]
        self.world, self.group = 1, None
        self._flat, self._graphs = None, {}

    def enable_data_parallel(self, world_size, group=None, reduce_all=False):
        self.world, self.group = int(world_size), group
        for m in self.models:
            m.engine.enable_data_parallel(world_size, group=group, reduce_all=reduce_all, overlap=False)
            m.engine.dp_defer = True      # this class reduces the gradients of all actions at once
        n = sum(m.engine.g_motion.g.numel() for m in self.models)
        self._flat = torch.empty(n, device=self.dev)

    def _allreduce_all(self):
        import torch.distributed as dist
        grads = [m.engine.g_motion.g for m in self.models]
        torch.cat(grads, out=self._flat)
        dist.all_reduce(self._flat, group=self.group)
        o = 0
        for g in grads:
            g.copy_(self._flat[o:o + g.numel()])
            o += g.numel()
        if self.models[0].engine.dp_reduce_all:
            small = [g.g for m in self.models for g in (m.engine.g_cam, m.engine.g_phase, m.engine.g_inst) if g is not None]
            flat = torch.cat(small)
            dist.all_reduce(flat, group=self.group)
            o = 0
            for g in small:
                g.copy_(flat[o:o + g.numel()])
                o += g.numel()

    def forward_backward(self, view_idx, frame_idx, update=True):
        """view_idx / frame_idx: one pair per action (device tensors for the no-copy path).  Returns the engines'
        (workspace, norm) pairs; the optimisers have NOT stepped yet."""
        cur = torch.cuda.current_stream(self.dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        outs = []
        for m, st, vi, fi in zip(self.models, self.streams, view_idx, frame_idx):
            st.wait_event(fork)
            with torch.cuda.stream(st):
                e = m.engine
                e.training = m.training or (update and m.VERSION >= 3)
                outs.append(e.forward_backward(vi, fi, update=update, mode='step'))
                e.training = False
            ev = torch.cuda.Event()
            ev.record(st)
            cur.wait_event(ev)
        if update and self.world > 1:
            self._allreduce_all()
        return outs

    def _fb_body(self):
        """every action's forward + backward on its own stream (and, on one GPU, its Adam)"""
        cur = torch.cuda.current_stream(self.dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        outs = []
        for m, st in zip(self.models, self.streams):
            st.wait_event(fork)
            with torch.cuda.stream(st):
                e = m.engine
                w = e.ws(self._n)
                e.training = m.training or m.VERSION >= 3
                outs.append(e.forward_backward(w['vi'], w['fi'], update=True, mode='step', _indices_loaded=True))
                if self.world == 1:
                    e.optimizer_step([o._g for o in m.optimizers])
                e.training = False
            ev = torch.cuda.Event()
            ev.record(st)
            cur.wait_event(ev)
        return outs

    def _adam_body(self):
        for m in self.models:
            m.engine.optimizer_step([o._g for o in m.optimizers])

    def _graphed(self, key, fn):
        ent = self._graphs.get(key)
        if ent is None:
            self._graphs[key] = 'warm'
            return fn()
        if ent == 'warm':
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.dev)
            with torch.cuda.graph(g):
                out = fn()
            ent = self._graphs[key] = (g, out)
        ent[0].replay()
        return ent[1]

    def train_step(self, view_idx, frame_idx):
        """forward + backward + Adam of every action as ONE CUDA-graph launch: the actions' kernels sit on parallel
        branches of the graph (one stream each at capture time).  Data parallel: graph 1 = every forward + backward, one
        EAGER all-reduce over all actions' shared gradients (NCCL stays outside the graphs), graph 2 = every Adam.  First
        call eager, second captures, later ones replay; the indices are copied into the engines' workspace slots outside
        the graphs.  Returns the engines' (workspace, norm) pairs."""
        self._n = int(view_idx[0].shape[0])
        for m, vi, fi in zip(self.models, view_idx, frame_idx):
            w = m.engine.ws(self._n)
            m.engine._flush_info()
            m.engine.step_tick += 1
            w['vi'].copy_(vi, non_blocking=True)
            w['fi'].copy_(fi, non_blocking=True)
            for o in m.optimizers:
                o.sync_hyper()
        key = (self._n, tuple((g.lr, g.wd) for m in self.models for g in m.engine.groups()))
        outs = self._graphed(('fb',) + key, self._fb_body)
        if self.world > 1:
            self._allreduce_all()
            self._graphed(('adam',) + key, self._adam_body)
        return outs

    def optimizer_step(self):
        for m in self.models:
            m._step_optimizers()

    def step(self, view_idx, frame_idx, update=True):
        """One iteration of every action model -> list of (loss_dict, info_dict), like `NemoV*.step`."""
        vi = [m._idx(v) for m, v in zip(self.models, view_idx)]
        fi = [m._idx(f) for m, f in zip(self.models, frame_idx)]
        outs = self.train_step(vi, fi) if update else self.forward_backward(vi, fi, update=False)
        host = self.losses_packed(outs)          # ONE device->host copy for all actions
        res = []
        for m, (w, norm), row in zip(self.models, outs, host):
            ld = m.engine.loss_dict_from_host(row, norm)
            res.append(({k: ld[k] for k in m._LOSS_KEYS}, _StepInfo(m.engine, w, m.points2d_gt_all)))
        return res

    def losses_packed(self, outs):
        """all actions' loss vectors in ONE device->host copy: (A, 9) = [acc(8) | total]"""
        vec = torch.stack([torch.cat([w['acc'], w['total']]) for w, _ in outs])
        if self.world > 1:
            torch.distributed.all_reduce(vec, group=self.group)
            vec /= self.world
        return vec.cpu().numpy().astype(np.float64)
