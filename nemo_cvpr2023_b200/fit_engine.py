"""Hand-scheduled forward + backward + Adam of one NeMo fitting iteration on one GPU.

This is the engine behind `nemo.neural_motion_model.NemoV1..V4.step / get_preds_batch / warmup /
opt_cam`.  It replaces the autograd graph of ~15.6k ATen calls the reference builds per step
 with ~90 launches of the kernels in libnemo_b200.so on preallocated buffers;
nothing here runs on the CPU and no per-vertex intermediate is written to HBM.

Reference semantics reproduced (paths under /root/reference, file nemo/neural_motion_model.py unless noted):
  get_preds_batch :3637-3672 (V4 :3968-4008), get_preds_given_phases_and_view :3733-3781 / :3911-3956,
  _get_smpl_given_poses :2963-2976, learned_camera_projection :3073-3124, keypoint_loss :2806-2843,
  per-view reduction :3551-3558, vposer_loss :2775-2804, gmm_prior_loss :2758-2773, 3D loss :3870-3882,
  instance loss :3864-3867, optimisers :3404-3431 / :3694-3721, warmup :3455-3509, opt_cam :2869-2906 / :4060-4149.
"""
from __future__ import annotations

import math
import weakref

import numpy as np
import torch

from . import assets
from ._lib import ADAM_OWNS_STEP, ADAM_SKIP, AdamSeg, adam_multi, lib, require_cuda
from .functional import KP_LOSSES, NV, RBF_BASES, VP, SMPLConstants, gemm, pose_blend, pose_blend_dgrad

LDY = 148      # row stride of the (N+1, 147) MLP output / its gradient: 16-byte rows for the TMA-fed GEMM
LDF = 208      # row stride of the (N, 207) pose feature, same reason
H_VP = 512     # VPoser hidden width
L_VP = 32      # VPoser latent size


class PendingLosses:
    """The losses of a step whose device->host copy may still be in flight (`NemoV*.step(..., sync=False)`)."""

    def __init__(self, host, event, norm, keys=None):
        self._host, self._event, self._norm, self._keys, self._out = host, event, norm, keys, None

    def done(self):
        return self._out is not None or self._event.query()

    def result(self):
        """-> the loss_dict `step` returns (numpy float32 scalars); waits for this step's copy if it has not landed"""
        if self._out is None:
            self._event.synchronize()
            ld = FitEngine.loss_dict_from_host(self._host.numpy().astype(np.float64), self._norm)
            self._out = ld if self._keys is None else {k: ld[k] for k in self._keys}
            self._host = self._event = None
        return self._out


class ParamGroup:
    """One optimiser of the reference = one flat fp32 buffer (+ grad, Adam moments, device step count)."""

    @staticmethod
    def _norm(specs):
        return [(sp[0], tuple(sp[1]), (sp[2] if len(sp) > 2 else (sp[1][-1] if len(sp[1]) else 1))) for sp in specs]

    @staticmethod
    def _sizes(specs):
        sizes = []
        for _, shape, ld in ParamGroup._norm(specs):
            k = int(np.prod(shape[:-1])) * ld if len(shape) >= 2 else int(np.prod(shape))
            sizes.append((k + 3) // 4 * 4)
        return sizes

    @staticmethod
    def flat_size(specs):
        return sum(ParamGroup._sizes(specs))

    def __init__(self, specs, device, lr, weight_decay=0.0, decoupled=False, storage=None):
        """specs: [(name, shape)] or [(name, shape, ld)].  Every tensor starts on a 16-byte boundary and a matrix may
        be stored with a padded row stride ld >= shape[1] (zero padding that Adam leaves at zero: zero gradient,
        zero weight), which is what the TMA-fed GEMM needs; `views` are the logical (unpadded) tensors, `padded` /
        `gpadded` the storage handed to the kernels.  storage = (p, g, m, v): zeroed flat slices of a buffer shared with
        the other groups of an engine (one allocation, one gradient fill per step)."""
        self.specs = self._norm(specs)
        sizes = self._sizes(specs)
        n = sum(sizes)
        if storage is None:
            storage = tuple(torch.zeros(n, device=device) for _ in range(4))
        self.p, self.g, self.m, self.v = storage
        assert all(t.numel() == n for t in storage)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=device)
        self.lr, self.wd, self.decoupled = float(lr), float(weight_decay), bool(decoupled)
        self.betas, self.eps = (0.9, 0.999), 1e-8
        self.views, self.gviews, self.offsets, self.padded, self.gpadded = {}, {}, {}, {}, {}
        o = 0
        for (name, shape, ld), size in zip(self.specs, sizes):
            if len(shape) >= 2:
                rows = int(np.prod(shape[:-1]))
                k = rows * ld
                self.padded[name] = self.p[o:o + k].view(rows, ld)
                self.gpadded[name] = self.g[o:o + k].view(rows, ld)
                self.views[name] = self.padded[name][:, :shape[-1]].view(*shape) if ld == shape[-1] \
                    else self.padded[name][:, :shape[-1]]
                self.gviews[name] = self.gpadded[name][:, :shape[-1]].view(*shape) if ld == shape[-1] \
                    else self.gpadded[name][:, :shape[-1]]
            else:
                k = int(np.prod(shape))
                self.views[name] = self.p[o:o + k].view(shape)
                self.gviews[name] = self.g[o:o + k].view(shape)
            self.offsets[name] = (o, k)
            o += size

    def set_lagging(self, names):
        """Tensors the reference's warm-up never reaches (grad None -> torch.optim.Adam keeps no state for them and
        their per-parameter step count lags, nemo/neural_motion_model.py:3455-3509): they get their own device step
        counter and can be skipped by `adam_step(skip_lag=True)`."""
        self.lag_names = [n for n in names if n in self.offsets]
        self.step_lag_dev = torch.zeros(1, dtype=torch.int32, device=self.p.device)

    def ranges(self):
        """[(offset, size, lagging)] covering the flat buffer, maximal runs of equal class."""
        lag = set(getattr(self, 'lag_names', ()))
        if not lag:
            return [(0, self.p.numel(), False)]
        names = [sp[0] for sp in self.specs]
        starts = [self.offsets[n][0] for n in names] + [self.p.numel()]
        out = []
        for i, n in enumerate(names):
            cls = n in lag
            if out and out[-1][2] == cls:
                out[-1] = (out[-1][0], starts[i + 1] - out[-1][0], cls)
            else:
                out.append((starts[i], starts[i + 1] - starts[i], cls))
        return out

    def segments(self, skip_lag=False):
        segs, owned = [], set()
        for o, k, lagging in self.ranges():
            step = self.step_lag_dev if lagging else self.step_dev
            flags = 0
            if step.data_ptr() not in owned:
                flags |= ADAM_OWNS_STEP
                owned.add(step.data_ptr())
            if lagging and skip_lag:
                flags |= ADAM_SKIP
            segs.append(AdamSeg(self.p.data_ptr() + 4 * o, self.g.data_ptr() + 4 * o, self.m.data_ptr() + 4 * o,
                                self.v.data_ptr() + 4 * o, k, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                                int(self.decoupled), step.data_ptr(), flags))
        return segs

    def adam_step(self, skip_lag=False):
        if not hasattr(self, 'ticket'):
            self.ticket = torch.zeros(1, dtype=torch.int32, device=self.p.device)
        adam_multi(self.segments(skip_lag), self.ticket)

    def fresh_alias(self):
        """A second optimiser over the SAME parameters and gradients with its own zeroed moments and step count (the
        reference's throw-away camera Adam, nemo/neural_motion_model.py:2869-2876)."""
        import copy
        a = copy.copy(self)
        a.m, a.v, a.step_dev = torch.zeros_like(self.m), torch.zeros_like(self.v), torch.zeros_like(self.step_dev)
        a.ticket = torch.zeros(1, dtype=torch.int32, device=self.p.device)
        return a

    def reset_state(self):
        self.m.zero_(), self.v.zero_(), self.step_dev.zero_()
        if hasattr(self, 'step_lag_dev'):
            self.step_lag_dev.zero_()


def motion_specs(h, in_dim, ldx, D):
    """Layout of the opt_motion buffer (MotionNet + RBF log_sigmas, nemo/neural_motion_model.py:3694-3711).  The two
    output heads are adjacent -- [rot_out.weight; linear_out.weight] is the (147, h) operand of ONE head GEMM, and
    [rot_out.bias; linear_out.bias] its bias -- and the first layer's rows are padded to 16 bytes for the TMA-fed GEMM,
    so the flat order is NOT torch's parameter order; FusedAdam maps between the two."""
    specs = [('learned_motion.net.net.0.weight', (h, in_dim), ldx), ('learned_motion.net.net.0.bias', (h,)),
             ('learned_motion.net.net.2.weight', (h, h)), ('learned_motion.net.net.2.bias', (h,)),
             ('learned_motion.net.net.4.weight', (h, h)), ('learned_motion.net.net.4.bias', (h,)),
             ('learned_motion.rot_out.weight', (144, h)), ('learned_motion.linear_out.weight', (3, h)),
             ('learned_motion.rot_out.bias', (144,)), ('learned_motion.linear_out.bias', (3,))]
    if D > 0:
        specs.append(('phase_rbf.log_sigmas', (D,)))
    return specs


LAGGING = ['learned_motion.linear_out.weight', 'learned_motion.linear_out.bias']


def fold_vposer(sd, device):
    """Eval-mode VPoser (human_body_prior/models/vposer_model.py:63-88, model_loader.py:83) as plain
    affine maps: BatchNorm1d running statistics are folded into the neighbouring Linear layers, and the
    activation-free chain Linear(512,512) -> Linear(512,512) -> {mu, logvar} is merged into one 512->64
    map (computed in float64).  Weights are frozen (model_loader.py:65-67) so this is done once."""
    g = lambda k: torch.as_tensor(np.asarray(sd[k]), dtype=torch.float64)
    eps = 1e-5
    s1 = g('encoder_net.1.weight') / torch.sqrt(g('encoder_net.1.running_var') + eps)
    o1 = g('encoder_net.1.bias') - g('encoder_net.1.running_mean') * s1
    W1, b1 = g('encoder_net.2.weight'), g('encoder_net.2.bias')
    W1f, b1f = W1 * s1[None], b1 + W1 @ o1
    s2 = g('encoder_net.4.weight') / torch.sqrt(g('encoder_net.4.running_var') + eps)
    o2 = g('encoder_net.4.bias') - g('encoder_net.4.running_mean') * s2
    W6, b6 = g('encoder_net.6.weight'), g('encoder_net.6.bias')
    W6f, b6f = W6 * s2[None], b6 + W6 @ o2
    W7, b7 = g('encoder_net.7.weight'), g('encoder_net.7.bias')
    Wh = torch.cat([g('encoder_net.8.mu.weight'), g('encoder_net.8.logvar.weight')], 0)
    bh = torch.cat([g('encoder_net.8.mu.bias'), g('encoder_net.8.logvar.bias')], 0)
    Whf = Wh @ W7 @ W6f
    bhf = Wh @ (W7 @ b6f + b7) + bh
    f = lambda t: t.float().to(device).contiguous()
    W1e = torch.zeros(W1f.shape[0], 72, dtype=torch.float64)
    W1e[:, 3:66] = W1f
    return dict(W1=f(W1f), W1e=f(W1e), b1=f(b1f), Wh=f(Whf), bh=f(bhf),
                Wd0=f(g('decoder_net.0.weight')), bd0=f(g('decoder_net.0.bias')),
                Wd3=f(g('decoder_net.3.weight')), bd3=f(g('decoder_net.3.bias')),
                Wd5=f(g('decoder_net.5.weight')), bd5=f(g('decoder_net.5.bias')))


_NVTX = bool(int(__import__('os').environ.get('NEMO_B200_NVTX', '0')))
_DP_DRY = bool(int(__import__('os').environ.get('NEMO_B200_DP_DRY', '0')))
# NEMO_B200_SIDE_SMALL=0 keeps the small independent kernels of a step (shape blend, GMM prior, 3D pose loss, VPoser-KL
# backward, bias column sums) on the main stream instead of the side stream (measurement switch; default: side stream)
_SIDE_SMALL = bool(int(__import__('os').environ.get('NEMO_B200_SIDE_SMALL', '1')))
# experimental (NEMO_B200_DP_GRAPH_NCCL=1): data-parallel steps as ONE CUDA graph with the bucketed NCCL all-reduces
# captured on the communication stream.  Measured on 2 GPUs: no gain (2.073 vs 2.076 ms/step -- the compute kernels are
# persistent, one CTA per SM with ~200 KB of shared memory, so an NCCL kernel only becomes resident between them and the
# ~50 us of the 9 MB all-reduce stay exposed either way), and the capture hung the 2-rank parity test once.  Default 0:
# two graphs around one eager all-reduce.
_DP_GRAPH_NCCL = bool(int(__import__('os').environ.get('NEMO_B200_DP_GRAPH_NCCL', '0')))
# NEMO_B200_DP_SYMM=1: the shared gradient is reduced through NVSwitch multicast (symmetric memory, multimem.ld_reduce /
# multimem.st) by torch's kernel inside the step graph instead of an eager NCCL all-reduce between two graphs (measured
# slower than NCCL).  NEMO_B200_DP_SYMM=2: this library's own peer-memory reduction (csrc/allreduce_p2p.cu: stage,
# reduce-scatter + all-gather through P2P loads / stores, copy out) inside the step graph.
_DP_SYMM = int(__import__('os').environ.get('NEMO_B200_DP_SYMM', '0'))


class _Range:
    """NVTX range around a phase of the step (NEMO_B200_NVTX=1): shows up in nsys / ncu --nvtx timelines."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if _NVTX:
            torch.cuda.nvtx.range_pop()


def _on_device(fn):
    """Run an engine method with the engine's device current (streams, events and the library's per-device state
    are created against the current device)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        with torch.cuda.device(self.dev):
            return fn(self, *a, **k)
    return wrapped


class FitEngine:
    def __init__(self, args, num_views, num_frames, smpl_consts: SMPLConstants, vposer_sd, gmm, data,
                 version=2, device='cuda', init_params=None, focal=assets.FOCAL_LENGTH):
        require_cuda()
        self.args, self.V, self.F, self.C_, self.version = args, int(num_views), int(num_frames), smpl_consts, version
        self.dev = torch.device(device)
        a = args
        self.h = int(a.h_dim)
        self.D = int(a.phase_rbf_dim) if version >= 2 else 0
        self.Cc = int(a.instance_code_size)
        self.K = int(a.monotonic_network_n_nodes)
        self.in_dim = (self.D if self.D > 0 else 1) + self.Cc
        self.ldx = (self.in_dim + 3) // 4 * 4      # row stride of X, dX and the first-layer weight
        self.basis = RBF_BASES.index(a.rbf_kernel) if self.D > 0 else 1
        self.focal = float(focal)
        self.IMG_D0, self.IMG_D1 = int(data['IMG_D0']), int(data['IMG_D1'])
        h, V, K = self.h, self.V, self.K
        # ---- parameters, grouped as the reference groups its optimisers
        # One allocation holds the parameters / gradients / moments of all optimisers, laid out
        # [cameras | motion MLP ... RBF log_sigmas | phase networks | instance codes].  Most gradients are OVERWRITTEN by
        # their producers every step (wgrad GEMM epilogues, column sums, cam_finish); only log_sigmas, the phase networks
        # and the instance codes are accumulated with atomics (embed_bwd) -- they sit at the end of the motion range and
        # right after it, so the per-step gradient fill is ONE small memset (`g_zero`) instead of one 9 MB fill per group.
        cam_specs = [('learned_cameras', (V, 9))]
        motion = motion_specs(h, self.in_dim, self.ldx, self.D)
        phase_specs = [('shifts', (V, K)), ('scales', (V, K))]
        inst_specs = [('learned_instance_code', (V, self.Cc))] if self.Cc else None
        sizes = [ParamGroup.flat_size(sp) if sp else 0 for sp in (cam_specs, motion, phase_specs, inst_specs)]
        flat = [torch.zeros(sum(sizes), device=self.dev) for _ in range(4)]
        offs = np.concatenate([[0], np.cumsum(sizes)])
        sto = lambda i: tuple(t[offs[i]:offs[i + 1]] for t in flat)
        self.g_cam = ParamGroup(cam_specs, self.dev, a.lr_camera, storage=sto(0))
        self.g_motion = ParamGroup(motion, self.dev, a.lr_human, a.wd_human, decoupled=(a.opt_human == 'adamw'),
                                   storage=sto(1))
        self.g_motion.set_lagging(LAGGING)
        self.g_phase = ParamGroup(phase_specs, self.dev, a.lr_phase, storage=sto(2))
        self.g_inst = ParamGroup(inst_specs, self.dev, a.lr_instance, storage=sto(3)) if self.Cc else None
        z0 = int(offs[1]) + (self.g_motion.offsets['phase_rbf.log_sigmas'][0] if self.D > 0 else sizes[1])
        self.g_zero = flat[1][z0:]              # gradients accumulated with atomics
        self.g_all = flat[1]
        self.betas = torch.zeros(1, 10, device=self.dev)
        self.betas_grad = torch.zeros(1, 10, device=self.dev)
        self.compute_betas_grad = False
        # the 2-D branch of a step (joint read-outs -> projection -> keypoint loss and its backward down to dR) is a chain
        # of small, latency-bound kernels that is independent of the prior branch (VPoser, two meshes, GMM) until the
        # gradients meet at the pose head: it runs on a side stream and fills the SMs the big kernels leave idle
        self.overlap_2d = True
        self._side = None
        self.centres = torch.linspace(0, 1, max(self.D, 1), device=self.dev)
        mv = self.g_motion.views
        o0 = self.g_motion.offsets['learned_motion.rot_out.weight'][0]
        self.Whead = self.g_motion.p[o0:o0 + 147 * h].view(147, h)
        self.gWhead = self.g_motion.g[o0:o0 + 147 * h].view(147, h)
        o1 = self.g_motion.offsets['learned_motion.rot_out.bias'][0]
        self.bhead = self.g_motion.p[o1:o1 + 147]
        self.gbhead = self.g_motion.g[o1:o1 + 147]
        if init_params is not None:
            self.load_params(init_params)
        # ---- frozen priors and data
        self.vp = fold_vposer(vposer_sd, self.dev)
        f32 = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32, device=self.dev).contiguous()
        self.gmm_means, self.gmm_prec, self.gmm_w = f32(gmm['means']), f32(gmm['precisions']), f32(gmm['nll_weights']).reshape(-1)
        self.gmm_psym = torch.empty(self.gmm_means.shape[0], 69, 72, device=self.dev)   # symmetrised once: frozen buffer
        lib.gmm_prepare(self.gmm_prec, self.gmm_means.shape[0], 69, self.gmm_psym)
        self.gt2d = f32(data['points2d_gt_all'])
        self.bbox = f32(data['gt_bbox_size'])
        self.hmr_theta = f32(data['hmr_theta']) if 'hmr_theta' in data else None
        self.hmr_mask = f32(data['hmr_mask']) if 'hmr_mask' in data else None
        self.phase_table = torch.linspace(0, 1, self.F, device=self.dev)
        # ---- read-outs the losses need (SURVEY.md section 8(a) "which joints the losses consume")
        jm = np.asarray(assets.JOINT_MAP_49)
        sel = list(range(25)) if version >= 4 else [38] + list(range(1, 25))
        raw = jm[sel]
        picks = sorted({int(r) - 24 for r in raw if 24 <= r < 45})
        extras = sorted({int(r) - 45 for r in raw if r >= 45})
        vids = np.asarray(assets.VERTEX_JOINT_IDS)[picks]
        self.ro = smpl_consts.make_readout(vids, smpl_consts.raw['J_regressor_extra'][extras].cpu() if extras
                                           else torch.zeros(0, NV))
        m25 = []
        for r in raw:
            r = int(r)
            m25.append(r if r < 24 else (24 + picks.index(r - 24) if r < 45 else 24 + len(picks) + extras.index(r - 45)))
        self.map25 = torch.tensor(m25, dtype=torch.int32, device=self.dev)
        self._ws = {}
        self.training = False
        self.world, self.loss_scale, self.dp_defer = 1, 1.0, False
        self.use_graph, self._graphs = True, {}
        self.step_tick = 0        # forward_backward calls so far (lazily read step outputs check it)

    def enable_data_parallel(self, world_size, group=None, reduce_all=False, overlap=True):
        """Shard by video instance: this rank owns its views' cameras, instance codes and
        phase networks; the shared motion MLP + RBF gradients are summed over ranks.  The global loss is the mean of
        the ranks' losses (equal shards), so local gradients are scaled by 1/world.

        overlap: the flat gradient is reduced in four buckets (head, layer 4, layer 2, layer 0 + RBF widths) on a
        communication stream as soon as the backward has produced each of them, so only the last, small bucket is
        exposed; the compute stream joins before Adam.  reduce_all: a view's frames are split over ranks (more GPUs
        than views), so its camera / phase / instance gradients are summed as well (one more small all-reduce)."""
        self.world, self.loss_scale = int(world_size), 1.0 / int(world_size)
        self.dp_group, self.dp_reduce_all, self.dp_overlap = group, bool(reduce_all), bool(overlap)
        self._dp_handles, self._comm = [], None
        off = self.g_motion.offsets
        o4, o2 = off['learned_motion.net.net.4.weight'][0], off['learned_motion.net.net.2.weight'][0]
        oh = off['learned_motion.rot_out.weight'][0]
        ot = off['phase_rbf.log_sigmas'][0] if self.D > 0 else self.g_motion.g.numel()
        self._dp_buckets = dict(head=(oh, ot), l4=(o4, oh), l2=(o2, o4), l0=(0, o2), tail=(ot, self.g_motion.g.numel()))
        self._symm = None
        if _DP_SYMM and self.world > 1 and not reduce_all:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else dist.group.WORLD
            n = (self.g_motion.g.numel() + 1023) // 1024 * 1024
            with torch.cuda.device(self.dev):
                buf = symm_mem.empty(n + 64, dtype=torch.float32, device=self.dev)   # + the signal area of the P2P kernels
                self._symm_hdl = symm_mem.rendezvous(buf, grp.group_name)
            buf.zero_()
            self._symm, self._symm_group = buf, grp.group_name
            if _DP_SYMM == 2:
                import ctypes
                hdl = self._symm_hdl
                assert hdl.world_size == self.world and self.g_motion.g.numel() % 4 == 0
                self._p2p_ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
                self._p2p = (ctypes.addressof(self._p2p_ptrs), int(hdl.rank), n)
                torch.cuda.synchronize(self.dev)
                dist.barrier(group=grp)          # every rank's signal area is zero before anybody signals

    def track_info(self, info):
        import weakref
        self._pending_info = weakref.ref(info)

    def _flush_info(self):
        """the previous step's lazily read outputs, if somebody still holds them, are copied out of the workspace now"""
        ref = getattr(self, '_pending_info', None)
        if ref is not None:
            self._pending_info = None
            obj = ref()
            if obj is not None:
                obj.materialize()

    def set_frame_window(self, frame_lo, total_frames):
        """This engine holds frames [frame_lo, frame_lo + F) of sequences that are `total_frames` long (data-parallel
        frame split): the raw phase of a frame is its position in the WHOLE sequence (:2978-2984)."""
        self.phase_table = torch.linspace(0, 1, int(total_frames), device=self.dev)[int(frame_lo):int(frame_lo) + self.F].contiguous()

    def _dp_bucket(self, name):
        """all-reduce one finished range of the shared gradient on the communication stream"""
        if self.world <= 1 or not self.dp_overlap or self.dp_defer:
            return
        import torch.distributed as dist
        lo, hi = self._dp_buckets[name]
        if hi <= lo:
            return
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=self.dev)
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ev)
            self._dp_handles.append(dist.all_reduce(self.g_motion.g[lo:hi], group=self.dp_group, async_op=True))

    def _symm_allreduce(self):
        """sum of the shared gradient over the ranks through the NVSwitch multicast address (graph-capturable)"""
        g, n = self.g_motion.g, self.g_motion.g.numel()
        if getattr(self, '_p2p', None) is not None:
            if not _DP_DRY:
                ptrs, rank, sig_off = self._p2p
                lib.p2p_allreduce(ptrs, rank, self.world, g, g, n, sig_off)
            return
        self._symm[:n].copy_(g)
        torch.ops.symm_mem.multimem_all_reduce_(self._symm[:(n + 1023) // 1024 * 1024], 'sum', self._symm_group)
        g.copy_(self._symm[:n])

    def _dp_finish(self, plain=False):
        import torch.distributed as dist
        if _DP_DRY:      # debug (NEMO_B200_DP_DRY=1): everything but the collective, to separate its cost from the rest
            return
        if self.dp_overlap and not plain:
            self._dp_bucket('tail')
        else:
            dist.all_reduce(self.g_motion.g, group=self.dp_group)
        if self.dp_reduce_all:
            small = [g.g for g in (self.g_cam, self.g_phase, self.g_inst) if g is not None]
            flat = torch.cat(small)
            dist.all_reduce(flat, group=self.dp_group)
            o = 0
            for g in small:
                g.copy_(flat[o:o + g.numel()])
                o += g.numel()
        for h in self._dp_handles:
            h.wait()          # the compute stream waits for the collective; the host does not block
        self._dp_handles = []

    # ------------------------------------------------------------------ parameters
    def named_params(self):
        out = dict(self.g_cam.views)
        out.update(self.g_motion.views)
        for v in range(self.V):
            out['phase_networks.%d.shifts' % v] = self.g_phase.views['shifts'][v]
            out['phase_networks.%d.scales' % v] = self.g_phase.views['scales'][v]
        if self.g_inst:
            out.update(self.g_inst.views)
        out['learned_betas'] = self.betas
        return out

    def named_grads(self):
        out = dict(self.g_cam.gviews)
        out.update(self.g_motion.gviews)
        for v in range(self.V):
            out['phase_networks.%d.shifts' % v] = self.g_phase.gviews['shifts'][v]
            out['phase_networks.%d.scales' % v] = self.g_phase.gviews['scales'][v]
        if self.g_inst:
            out.update(self.g_inst.gviews)
        out['learned_betas'] = self.betas_grad
        return out

    def load_params(self, params):
        mine = self.named_params()
        for k, v in params.items():
            if k == 'phase_rbf.centres':
                self.centres.copy_(torch.as_tensor(np.asarray(v), dtype=torch.float32).reshape(-1))
            elif k in mine:
                mine[k].copy_(torch.as_tensor(np.asarray(v), dtype=torch.float32).reshape(mine[k].shape))

    def groups(self):
        return [g for g in (self.g_cam, self.g_motion, self.g_phase, self.g_inst) if g is not None]

    # ------------------------------------------------------------------ workspace
    def ws(self, N):
        if N in self._ws:
            return self._ws[N]
        d, h = self.dev, self.h
        Np = N + 1
        z = lambda *s: torch.zeros(*s, device=d)
        nxr = self.ro['np'] + self.ro['nx']
        ldz = self.ro['ldz']
        w = dict(
            raw=z(N), X=z(Np, self.ldx), warped=z(N), yn=z(self.V, 3),
            H1=z(Np, h), H2=z(Np, h), H3=z(Np, h), Y=z(Np, LDY),
            R=z(N, 24, 9), aa=z(N, 72), trans=z(N, 3), F=z(N, LDF), F_m=z(N, LDF), dF_m=z(N, LDF), dF_o=z(N, LDF),
            PO_m=z(N, 3 * VP), PO_o=z(N, 3 * VP), PO_r=z(N, 3 * VP), dPO=z(N, 3 * VP),
            A_m=z(N, 288), A_o=z(N, 288), A_r=z(N, 288), JP=z(N, 72), JPs=z(N, 72),
            R_o=z(N, 24, 9), R_r=z(N, 24, 9),
            JX=z(N, max(nxr, 1), 3), Z=z(N, ldz), dZ=z(N, ldz), zb=z(ldz), dzb=z(ldz), j25=z(N, 25, 3), p2d=z(N, 25, 2), loss_all=z(N, 25, 2),
            E1=z(N, H_VP), head=z(N, 2 * L_VP), dhead=z(N, 2 * L_VP), D1=z(N, H_VP), D2=z(N, H_VP), D3=z(N, 126),
            Rdec=z(N * 21, 9), recon=z(N, 69), signs=torch.zeros(N, 3, 2, VP // 32, dtype=torch.int32, device=d),
            ll=z(N), terms=z(8), wts=z(8), total=z(1),
            d_aa=z(N, 72), dR=z(N, 24, 9), dR_o=z(N, 24, 9), dA=z(N, 288), dAs=z(N, 288), dJP=z(N, 72), dJX=z(N, max(nxr, 1), 3),
            dj25=z(N, 25, 3), dtr=z(N, 3), dY=z(Np, LDY), dHa=z(Np, h), dHb=z(Np, h), dX=z(Np, self.ldx),
            dE1=z(N, H_VP), gy01=z(self.V, 2), dJ=z(72), dJ_m=z(72), dbase=z(3 * VP),
            base=z(3, VP), J=z(72), vi=torch.zeros(N, dtype=torch.long, device=d),
            fi=torch.zeros(N, dtype=torch.long, device=d), noise=z(N, max(self.Cc, 1)),
        )
        # per-step accumulators (loss terms, per-view sums, camera gradient sums): one buffer, one fill per step
        V = self.V
        w['zbuf'] = z(8 + 14 * V + 4)
        w['acc'], w['view_acc'] = w['zbuf'][:8], w['zbuf'][8:8 + 2 * V].view(V, 2)
        w['dcam'], w['dyscr'] = w['zbuf'][8 + 2 * V:8 + 14 * V].view(V, 12), w['zbuf'][8 + 14 * V:]
        self._ws[N] = w
        return w

    # ------------------------------------------------------------------ forward pieces
    def _shape(self, w):
        """v_shaped and rest joints for the shared betas (blend_shapes + vertices2joints, lbs.py:257-300)."""
        C = self.C_
        gemm(self.betas, C.Sp, w['base'], 1, 3 * VP, 10, 10, 3 * VP, 3 * VP, bias=C.vt.view(-1))
        gemm(self.betas, C.JSt, w['J'], 1, 72, 10, 10, 72, 72, bias=C.Jt)
        ro = self.ro
        gemm(self.betas, ro['MZ'][207:], w['zb'], 1, ro['ldz'], 10, 10, ro['ldz'], ro['ldz'], bias=ro['Z0'])

    def _motion_forward(self, w, N, view_idx, raw_phase, use_noise):
        mv, gp, C = self.g_motion.views, self.g_phase.views, self.Cc
        Np, h = N + 1, self.h
        codes = self.g_inst.views['learned_instance_code'] if self.g_inst else None
        noise = None
        if use_noise:
            w['noise'].normal_().mul_(float(self.args.code_noise))
            noise = w['noise']
        lib.embed_fwd(raw_phase, view_idx, gp['shifts'], gp['scales'], self.K, self.centres,
                      mv['phase_rbf.log_sigmas'] if self.D > 0 else None, self.D, self.basis, codes, C, noise,
                      w['X'], self.ldx, w['warped'], w['yn'], N, self.V)
        W = lambda i: mv['learned_motion.net.net.%d.weight' % i]
        b = lambda i: mv['learned_motion.net.net.%d.bias' % i]
        gemm(w['X'], self.g_motion.padded['learned_motion.net.net.0.weight'], w['H1'], Np, h, self.in_dim, self.ldx,
             self.ldx, h, transB=True, bias=b(0), act=1)
        gemm(w['H1'], W(2), w['H2'], Np, h, h, h, h, h, transB=True, bias=b(2), act=1)
        # FCNN's last Linear followed by MotionNet.relu (:132)
        gemm(w['H2'], W(4), w['H3'], Np, h, h, h, h, h, transB=True, bias=b(4), act=1)
        gemm(w['H3'], self.Whead, w['Y'], Np, 147, h, h, h, LDY, transB=True, bias=self.bhead)
        lib.rot6d_fwd(w['Y'], LDY, w['R'], 216, 24, N * 24)
        lib.rotmat2aa_fwd(w['R'], 9, w['aa'], 3, 1, N * 24, 1)
        # trans - trans_0 (:3755-3766): row N of Y is the phase-0 / zero-code sample
        torch.sub(w['Y'][:N, 144:147], w['Y'][N:Np, 144:147], out=w['trans'])

    def _lbs_forward(self, w, N, R, PO, A, JP, aa=None, lda=0):
        """(Rodrigues ->) pose feature + kinematic chain in one launch (lbs.py:221-235, :337-404), then the pose blend
        shapes (planar).  aa: the body pose as axis-angle rows; R then RECEIVES the rotations (root copied from the
        main mesh's), as the two extra SMPL calls of vposer_loss build them (:2787-2793)."""
        C = self.C_
        if aa is None:
            lib.pose_chain_fwd(R, None, 0, 0, None, w['J'], 0, C.parents, None, w['F'], LDF, A, JP, N)
        else:
            lib.pose_chain_fwd(None, aa, lda, 0, w['R'], w['J'], 0, C.parents, R, w['F'], LDF, A, JP, N)
        pose_blend(w['F'], C, PO, N, LDF)

    def _joints_forward(self, w, N):
        """The main mesh of a step is consumed only through 25 joints: kinematic joints come
        from the chain, picked vertices and the regressed head joint from the read-out operators (csrc/readout.cu),
        which contract the pose feature with pre-reduced blend shapes -- identical arithmetic up to fp32 summation
        order, without the 6890-vertex mesh, its 20736-wide pose-blend GEMM or the skinning kernel."""
        C, ro = self.C_, self.ro
        # own feature buffer: the prior branch rewrites w['F'] concurrently
        lib.pose_chain_fwd(w['R'], None, 0, 0, None, w['J'], 0, C.parents, None, w['F_m'], LDF, w['A_m'], w['JP'], N)
        if ro['np'] + ro['nx']:
            gemm(w['F_m'], ro['MZ'], w['Z'], N, ro['ldz'], 207, LDF, ro['ldz'], ro['ldz'], bias=w['zb'])
            lib.readout_fwd(w['A_m'], w['Z'], ro['ldz'], ro['Wp'], ro['np'], ro['Sj'], ro['nx'], w['JX'], N)

    # ------------------------------------------------------------------ the step
    @_on_device
    def forward_backward(self, view_idx, frame_idx, update=True, phases=None, mode='step', detach_pose=False,
                         _indices_loaded=False):
        """mode: 'step' (all losses), 'warmup' (3D pose loss only), 'opt_cam' (V4: 2D + 3D, pose detached),
        'cam_fit' (MultiViewModel.opt_cam: plain-mean reprojection loss only)."""
        a, C, ro = self.args, self.C_, self.ro
        N = int(view_idx.shape[0])
        w = self.ws(N)
        Np, h = N + 1, self.h
        if not _indices_loaded:
            self._flush_info()
        self.step_tick += 1
        if not _indices_loaded:
            w['vi'].copy_(view_idx, non_blocking=True)
            w['fi'].copy_(frame_idx, non_blocking=True)
        vi, fi = w['vi'], w['fi']
        if phases is None:
            torch.index_select(self.phase_table, 0, fi, out=w['raw'])
        else:
            w['raw'].copy_(phases)
        w['zbuf'].zero_()
        use_noise = self.version >= 3 and self.training and a.code_noise > 0 and self.Cc > 0
        full = mode == 'step'
        need_2d = mode in ('step', 'opt_cam', 'cam_fit')
        # fork: with both branches present (a full step that updates), the 2-D branch goes to the side stream
        overlap = (self.overlap_2d and need_2d and mode == 'step' and update and not self.compute_betas_grad
                   and lib.profile is None)
        # ... and so do the small kernels that depend on nothing but the pose (or a finished gradient): on the main
        # stream each of them is 7-55 us of a nearly empty GPU on the critical path (profiles/r02_launches_v3.csv).
        # Not under the eager bucketed all-reduce, whose buckets follow the bias sums on the main stream.
        side_small = overlap and _SIDE_SMALL and (self.world == 1 or self.dp_defer)
        side = None
        if overlap:
            if self._side is None:
                self._side = (torch.cuda.Stream(device=self.dev), torch.cuda.Event(), torch.cuda.Event())
            side, ev_fork, ev_join = self._side
        main = torch.cuda.current_stream()

        def on_side(fn):
            """queue fn on the side stream behind everything queued on the main stream so far -> its completion event"""
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                fn()
                done = torch.cuda.Event()
                done.record(side)
            return done

        ev_shape = None
        if side_small:
            ev_shape = on_side(lambda: self._shape(w))
        with _Range('nemo.motion_forward'):
            self._motion_forward(w, N, vi, w['raw'], use_noise)
            if not side_small:
                self._shape(w)
        mode_id = KP_LOSSES.index(a.loss)
        red = 1 if mode == 'cam_fit' else 0
        cx, cy = float(self.IMG_D0 // 2), float(self.IMG_D1 // 2)   # (:3104-3106) height/2 is used as cx
        ls = self.loss_scale
        acc = w['acc']   # 0 l1_sum, 1 kl_sum, 2 loss_3d, 3 gmm_sum, 4 kp, 5 instance
        dR = w['dR']

        def branch_2d_forward():
            self._joints_forward(w, N)
            lib.joints_gather_fwd(w['JP'], w['JX'], ro['np'] + ro['nx'], ro['np'], self.map25, 25, w['trans'], None,
                                  w['j25'], N)
            lib.kp_loss_fwd(w['j25'], vi, fi, self.g_cam.views['learned_cameras'], self.gt2d, self.bbox, self.focal,
                            cx, cy, 100.0, mode_id, red, w['p2d'], w['loss_all'], w['view_acc'], N, 25, self.V, self.F)

        def branch_2d_backward():
            lib.kp_loss_bwd(w['j25'], vi, fi, self.g_cam.views['learned_cameras'], self.gt2d, self.bbox,
                            self.focal, cx, cy, 100.0, mode_id, red, w['view_acc'], ls, w['dj25'], w['dcam'],
                            acc[4:5], N, 25, self.V, self.F)
            lib.cam_finish(self.g_cam.views['learned_cameras'], w['dcam'], self.g_cam.gviews['learned_cameras'],
                           0, self.V)
            lib.joints_gather_bwd(w['dj25'], ro['np'] + ro['nx'], ro['np'], self.map25, 25, None, w['dJP'],
                                  w['dJX'], w['dtr'], N)
            nq = ro['np'] + ro['nx']
            if nq:
                lib.readout_bwd(w['A_m'], w['Z'], ro['ldz'], ro['Wp'], ro['np'], ro['Sj'], ro['nx'], w['dJX'],
                                w['dA'], 0, w['dZ'], N)
            # the read-outs' gradient with respect to the pose feature first (its own aligned rows: tensor-core GEMM with
            # the TMA epilogue), then the chain backward adds it into dR[:, 1:] while it writes dR.  V4 opt_cam
            # (:4060-4149) detaches the body rotmats: the gradient reaches the MLP through the global orientation and
            # the translation only.
            dF = None
            if nq and not detach_pose:
                dF = w['dF_m']
                gemm(w['dZ'], ro['MZ'], dF, N, 207, ro['ldz'], ro['ldz'], ro['ldz'], LDF, transB=True)
            lib.chain_bwd2(w['R'], w['J'], 0, C.parents, w['dA'] if nq else None, None, w['dJP'], dF, LDF, dR,
                           w['dJ_m'].zero_() if self.compute_betas_grad else None, N)
            if detach_pose:
                dR[:, 1:].zero_()
            if self.compute_betas_grad:
                w['dJ'].copy_(w['dJ_m'])
                self._betas_grad(w, N, first=True, mesh=False)

        w_vp = float(a.weight_vp_loss) if full else 0.0
        w_kl = float(a.weight_vp_z_loss) if full else 0.0
        w_gmm = float(a.weight_gmm_loss) if full else 0.0
        w_3d = float(getattr(a, 'weight_3d_loss', 0.0)) if (self.version >= 3 and mode in ('step', 'opt_cam')) else 0.0
        if mode == 'warmup':
            w_3d = 1.0
        w_inst = float(getattr(a, 'weight_instance_loss', 0.0)) if (self.version >= 3 and full and self.Cc) else 0.0

        def pose_priors():
            """GMM prior (:2758-2773: value and scaled gradient in one pass) and the 3D pose loss (:3870-3882): both read
            the predicted axis-angle pose only and ADD their gradient into d_aa (zeroed before)"""
            if full:
                lib.gmm_prior_sym(w['aa'].view(-1)[3:], 72, self.gmm_means, self.gmm_psym, self.gmm_w,
                                  self.gmm_means.shape[0], 69, w['ll'], ls * w_gmm / N,
                                  w['d_aa'].view(-1)[3:] if update else None, 72, 1, N)
                lib.colsum(w['ll'], acc[3:4], N, 1, 1, 1)
            if w_3d or (self.version >= 3 and full and getattr(a, 'weight_3d_loss', 0.0)):
                lib.pose3d_loss(w['aa'].view(-1)[3:], 72, vi, fi, self.hmr_theta, self.hmr_mask, 100.0, ls * w_3d,
                                acc[2:3], w['d_aa'].view(-1)[3:] if update else None, 72, 1, N, 69, self.F)

        def kl_backward():
            vp = self.vp
            gemm(w['dhead'], vp['Wh'], w['dE1'], N, H_VP, 2 * L_VP, 2 * L_VP, H_VP, H_VP, gate=w['E1'], ldg=H_VP,
                 gate_slope=0.01)
            gemm(w['dE1'], vp['W1e'], w['d_aa'], N, 72, H_VP, H_VP, 72, 72, accumulate=True)

        if update:
            self.g_zero.zero_()
        w['d_aa'].zero_()
        ev_daa = None        # side stream: every addition to d_aa queued there has been made
        if overlap:
            ev_fork.record()
            side.wait_event(ev_fork)
            with torch.cuda.stream(side):
                if side_small:
                    pose_priors()
                branch_2d_forward()
                branch_2d_backward()
                ev_join.record(side)
        elif need_2d:
            branch_2d_forward()
        if full:
            vp = self.vp
            # VPoser encode (mean) -> decode (:2779-2784)
            # W1e is the folded first encoder layer widened to all 72 axis-angle columns (zero columns for the global
            # orientation and the two hand joints), so the GEMM reads the 16-byte aligned (N,72) buffer directly
            gemm(w['aa'], vp['W1e'], w['E1'], N, H_VP, 72, 72, 72, H_VP, transB=True, bias=vp['b1'], act=2)
            gemm(w['E1'], vp['Wh'], w['head'], N, 2 * L_VP, H_VP, H_VP, H_VP, 2 * L_VP, transB=True, bias=vp['bh'])
            lib.vposer_kl(w['head'], L_VP, None, acc[1:2], ls * w_kl / N, w['dhead'], N)
            gemm(w['head'], vp['Wd0'], w['D1'], N, H_VP, L_VP, 2 * L_VP, L_VP, H_VP, transB=True, bias=vp['bd0'], act=2)
            gemm(w['D1'], vp['Wd3'], w['D2'], N, H_VP, H_VP, H_VP, H_VP, H_VP, transB=True, bias=vp['bd3'], act=2)
            gemm(w['D2'], vp['Wd5'], w['D3'], N, 126, H_VP, H_VP, H_VP, 126, transB=True, bias=vp['bd5'])
            lib.rot6d_fwd(w['D3'], 6, w['Rdec'], 9, 1, N * 21)
            lib.rotmat2aa_fwd(w['Rdec'], 189, w['recon'], 69, 21, N * 21, 0)
            w['recon'][:, 63:].copy_(w['aa'][:, 66:72])
            # the KL gradient (two small GEMMs that add into d_aa) needs nothing else: side stream, behind the pose priors
            if side_small and update and w_kl:
                ev_daa = on_side(kl_backward)
            # the two extra SMPL forwards of vposer_loss (:2787-2793): quaternion Rodrigues, shared global orient
            if ev_shape is not None:
                main.wait_event(ev_shape)
            self._lbs_forward(w, N, w['R_o'], w['PO_o'], w['A_o'], w['JPs'], aa=w['aa'].view(-1)[3:], lda=72)
            self._lbs_forward(w, N, w['R_r'], w['PO_r'], w['A_r'], w['JPs'], aa=w['recon'], lda=69)
            lib.skin_l1_fwd(w['A_o'], w['PO_o'], w['A_r'], w['PO_r'], w['base'], C.Wt, w['signs'], acc[0:1], N)
        if not side_small:
            pose_priors()
        if w_inst:
            code = self.g_inst.views['learned_instance_code']
            acc[5:6].copy_((code * code).mean().reshape(1))

        # ---------------------------------------------------------------- backward
        if update:
            if need_2d:
                if not overlap:
                    branch_2d_backward()
            else:
                dR.zero_()
                w['dtr'].zero_()
            if full and w_kl and ev_daa is None:
                kl_backward()
            if full and w_vp:
                ss = ls * w_vp / (N * NV * 3.0)
                lib.skin_bwd(w['A_o'], w['PO_o'], w['base'], C.Wt, None, w['signs'], ss, w['dPO'], w['dAs'], N)
                # |dPO| = |R^T dv| <= 3 |ss| (dv = +-ss or 0, blended rotation entries <= 1): known range, fp16-pair kernel
                pose_blend_dgrad(w['dPO'], C, w['dF_o'], N, LDF, 4.0 * abs(ss))
                lib.chain_bwd2(w['R_o'], w['J'], 0, C.parents, w['dAs'], None, None, w['dF_o'], LDF, w['dR_o'],
                               w['dJ'].zero_() if self.compute_betas_grad else None, N)
                if self.compute_betas_grad:
                    self._betas_grad(w, N, first=not need_2d, mesh=True)
                # d_aa: the side stream's additions (pose priors, KL) come before the main stream's
                if ev_daa is not None:
                    main.wait_event(ev_daa)
                elif side_small:
                    main.wait_event(ev_join)
                lib.rodrigues_bwd(w['aa'].view(-1)[3:], 72, w['dR_o'].view(-1)[9:], 216, w['d_aa'].view(-1)[3:], 72,
                                  23, 1, N * 23, 0)
                if overlap:
                    main.wait_event(ev_join)
                dR[:, 0].add_(w['dR_o'][:, 0])
            elif overlap:
                main.wait_event(ev_join)
                if ev_daa is not None:
                    main.wait_event(ev_daa)
            # axis-angle outputs of the pose head -> rotation matrices -> 6D -> MLP output
            lib.rotmat2aa_bwd(w['R'], 9, w['d_aa'], 3, dR, 9, 1, 1, N * 24, 1)
            lib.rot6d_bwd(w['Y'], LDY, dR, 216, w['dY'], LDY, 24, 0, N * 24)
            # translation columns, pad column and the phase-0 row of dY (trans - trans_0, :3755-3766) in one launch
            lib.dy_tail(w['dtr'] if need_2d else None, w['dY'], LDY, N, w['dyscr'])
            with _Range('nemo.motion_backward'):
                self._motion_backward(w, N, vi, on_side if side_small else None)
            if w_inst:
                self.g_inst.gviews['learned_instance_code'].add_(
                    self.g_inst.views['learned_instance_code'], alpha=ls * 2.0 * w_inst / (self.V * self.Cc))
            if self.world > 1 and not self.dp_defer:
                self._dp_finish()
        elif need_2d:
            lib.kp_loss_bwd(w['j25'], vi, fi, self.g_cam.views['learned_cameras'], self.gt2d, self.bbox, self.focal, cx,
                            cy, 100.0, mode_id, red, w['view_acc'], 0.0, w['dj25'], None, acc[4:5], N, 25, self.V, self.F)
        # ---------------------------------------------------------------- losses -> one packed vector
        norm = np.asarray([1.0 / (N * NV * 3.0), 1.0 / N, 1.0, 1.0 / N, 1.0, 1.0, 0.0, 0.0])
        wkey = (w_vp, w_kl, w_3d, w_gmm, need_2d, w_inst)
        cache = w.setdefault('wts_cache', {})
        if wkey not in cache:
            wt = np.asarray([w_vp, w_kl, w_3d, w_gmm, 1.0 if need_2d else 0.0, w_inst, 0.0, 0.0]) * norm
            cache[wkey] = torch.tensor(wt, dtype=torch.float32, device=self.dev)
        lib.combine_losses(acc, cache[wkey], 8, w['total'])
        return w, norm

    def _betas_grad(self, w, N, first, mesh):
        """d learned_betas: the reference accumulates it on every backward although no optimiser owns it.
        mesh=True: through the vertices (column sum of dPO); mesh=False: through the joint read-outs (dZ)."""
        C, ro = self.C_, self.ro
        if mesh:
            lib.colsum(w['dPO'], w['dbase'], N, 3 * VP, 3 * VP, 0)
            gemm(w['dbase'], C.Sp, self.betas_grad, 1, 10, 3 * VP, 3 * VP, 3 * VP, 10, transB=True, accumulate=True)
        elif ro['np'] + ro['nx']:
            lib.colsum(w['dZ'], w['dzb'], N, ro['ldz'], ro['ldz'], 0)
            gemm(w['dzb'], ro['MZ'][207:], self.betas_grad, 1, 10, ro['ldz'], ro['ldz'], ro['ldz'], 10, transB=True,
                 accumulate=True)
        gemm(w['dJ'], C.JSt, self.betas_grad, 1, 10, 72, 72, 72, 10, transB=True, accumulate=True)

    def _motion_backward(self, w, N, vi, on_side=None):
        """on_side: queue a bias column sum on the side stream behind the GEMM that produced its input (it returns the
        completion event); None: on the main stream, where the eager data-parallel buckets follow it."""
        mv, gv = self.g_motion.views, self.g_motion.gviews
        Np, h, ind = N + 1, self.h, self.in_dim
        W = lambda i: mv['learned_motion.net.net.%d.weight' % i]
        gW = lambda i: gv['learned_motion.net.net.%d.weight' % i]
        gb = lambda i: gv['learned_motion.net.net.%d.bias' % i]
        dY = w['dY']
        main = torch.cuda.current_stream()

        def bias_sum(X, out, ncol, ld):
            if on_side is None:
                lib.colsum(X, out, Np, ncol, ld, 0)
                return None
            return on_side(lambda: lib.colsum(X, out, Np, ncol, ld, 0))

        gemm(dY, w['H3'], self.gWhead, 147, h, Np, LDY, h, h, transA=True)
        # d/d linear_out.bias of (trans - trans_0) cancels exactly (row N of dY is minus the column sum of the rows above
        # it): only the 144 rot_out columns are summed, the three linear_out.bias entries are never written and stay an
        # exact 0 like the reference's -- not summation-order noise that Adam would amplify into lr-sized steps
        bias_sum(dY, self.gbhead, 144, LDY)
        self._dp_bucket('head')
        gemm(dY, self.Whead, w['dHa'], Np, h, 147, LDY, h, h, gate=w['H3'], ldg=h)
        gemm(w['dHa'], w['H2'], gW(4), h, h, Np, h, h, h, transA=True)
        ev_a = bias_sum(w['dHa'], gb(4), h, h)
        self._dp_bucket('l4')
        gemm(w['dHa'], W(4), w['dHb'], Np, h, h, h, h, h, gate=w['H2'], ldg=h)
        gemm(w['dHb'], w['H1'], gW(2), h, h, Np, h, h, h, transA=True)
        bias_sum(w['dHb'], gb(2), h, h)
        self._dp_bucket('l2')
        if ev_a is not None:
            main.wait_event(ev_a)      # dHa is rewritten next: its column sum on the side stream must have read it
        gemm(w['dHb'], W(2), w['dHa'], Np, h, h, h, h, h, gate=w['H1'], ldg=h)
        gemm(w['dHa'], w['X'], self.g_motion.gpadded['learned_motion.net.net.0.weight'], h, ind, Np, h, self.ldx,
             self.ldx, transA=True)
        ev_last = bias_sum(w['dHa'], gb(0), h, h)
        self._dp_bucket('l0')
        gemm(w['dHa'], self.g_motion.padded['learned_motion.net.net.0.weight'], w['dX'], Np, ind, h, h, self.ldx,
             self.ldx)
        gp, gg = self.g_phase.views, self.g_phase.gviews
        lib.embed_bwd(w['raw'], vi, gp['shifts'], gp['scales'], self.K, self.centres,
                      mv['phase_rbf.log_sigmas'] if self.D > 0 else None, self.D, self.basis, self.Cc, w['warped'],
                      w['yn'], w['dX'], self.ldx, gv['phase_rbf.log_sigmas'] if self.D > 0 else None, gg['shifts'],
                      gg['scales'], self.g_inst.gviews['learned_instance_code'] if self.g_inst else None, w['gy01'], N,
                      self.V)
        if ev_last is not None:
            main.wait_event(ev_last)   # join: the side stream is in order, its last column sum is behind all the others

    @_on_device
    def step_graph(self, view_idx, frame_idx, mode='step', detach_pose=False, which=None, skip_lag=False):
        """forward + backward + every optimiser as ONE CUDA-graph launch: ~90 kernel launches and
        the ATen fills / copies between them collapse into a single cudaGraphLaunch, which removes the host-side launch
        cost that separates the end-to-end rate from the device-timed one.  The first call with a new key (batch size,
        mode, training flag, optimiser hyper-parameters -- the learning rates are baked into the launch parameters, a
        plateau scheduler that changes one simply causes a re-capture) runs eagerly, the second captures, later ones
        replay.  The indices are copied into their fixed workspace slots outside the graph.  Under data parallelism NCCL
        stays outside the graphs: graph 1 = forward + backward, ONE eager all-reduce of the shared gradient, graph 2 = the
        optimisers (the bucketed, overlapped all-reduce belongs to the eager `forward_backward`;
        NEMO_B200_DP_GRAPH_NCCL=1 captures it instead -- experimental, see the top of the file).  Not used when
        per-entry-point profiling is on."""
        N = int(view_idx.shape[0])
        w = self.ws(N)
        self._flush_info()
        w['vi'].copy_(view_idx, non_blocking=True)
        w['fi'].copy_(frame_idx, non_blocking=True)
        groups = which or self.groups()
        key = (N, mode, bool(detach_pose), bool(self.training), bool(skip_lag), bool(self.compute_betas_grad),
               tuple((id(g), g.lr, g.wd, g.betas, g.eps, g.decoupled) for g in groups))

        symm = self.world > 1 and getattr(self, '_symm', None) is not None and not self.dp_defer
        dp = self.world > 1 and not symm and not (_DP_GRAPH_NCCL and self.dp_overlap and not self.dp_defer)

        def body():
            # data parallel: the graph stops before the all-reduce (NCCL stays OUTSIDE the captured graphs: one eager
            # collective between the two replays; a second graph holds the optimisers)
            saved = self.dp_defer
            self.dp_defer = self.dp_defer or dp or symm
            try:
                out = self.forward_backward(w['vi'], w['fi'], update=True, mode=mode, detach_pose=detach_pose,
                                            _indices_loaded=True)
            finally:
                self.dp_defer = saved
            if symm:
                self._symm_allreduce()
            if not dp:      # one GPU, or the collectives are part of the graph (forward_backward has joined them)
                self.optimizer_step(groups, skip_lag=skip_lag)
            return out

        def tail():
            if dp and not self.dp_defer:
                self._dp_finish(plain=True)
                self._adam_graph(groups, skip_lag, key)

        if lib.profile is not None or not self.use_graph:
            out = body()
            tail()
            return out
        ent = self._graphs.get(key)
        if ent is None:                       # first sight of this key: the eager run doubles as the warm-up
            self._graphs[key] = 'warm'
            out = body()
            tail()
            return out
        if ent == 'warm':
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.dev)
            with torch.cuda.graph(g):
                out = body()
            ent = self._graphs[key] = (g, out)
        ent[0].replay()
        self.step_tick += 1
        tail()
        return ent[1]

    def _adam_graph(self, groups, skip_lag, key):
        """the optimisers of a data-parallel step as their own small graph (after the eager all-reduce)"""
        k = ('adam',) + key
        ent = self._graphs.get(k)
        if ent is None or not self.use_graph:
            self._graphs[k] = 'warm'
            return self.optimizer_step(groups, skip_lag=skip_lag)
        if ent == 'warm':
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.dev)
            with torch.cuda.graph(g):
                self.optimizer_step(groups, skip_lag=skip_lag)
            ent = self._graphs[k] = (g, None)
        ent[0].replay()

    def drop_graphs(self, group):
        """forget the captured steps that reference `group`'s buffers"""
        self._graphs = {k: v for k, v in self._graphs.items() if not any(t[0] == id(group) for t in k[-1])}

    @_on_device
    def optimizer_step(self, which=None, skip_lag=False):
        """Adam / AdamW of every optimiser in ONE launch (nemo_b200_adam_multi)."""
        segs = []
        for g in (which or self.groups()):
            segs += g.segments(skip_lag)
        if not hasattr(self, '_adam_ticket'):
            self._adam_ticket = torch.zeros(1, dtype=torch.int32, device=self.dev)
        with _Range('nemo.adam'):
            adam_multi(segs, self._adam_ticket)

    # ------------------------------------------------------------------ read-backs
    def loss_dict(self, w, norm):
        """One packed device->host copy of the step's scalars (the reference does 5-7, :3564-3584)."""
        vec = torch.cat([w['acc'], w['total']])
        if self.world > 1:
            # global losses (the mean of the ranks' equal shards): what the reference would print on one GPU and what
            # the plateau schedulers must see so that every rank takes the same decision
            torch.distributed.all_reduce(vec, group=self.dp_group)
            vec /= self.world
        return self.loss_dict_from_host(vec.cpu().numpy().astype(np.float64), norm)

    def loss_dict_async(self, w, norm, keys=None):
        """The same read-back without blocking the host: the packed vector goes to a pinned slot with an asynchronous copy
        and the caller gets a `PendingLosses` whose `result()` waits for THAT copy only -- a fit loop that reads the
        losses of step i while step i + 1 runs never leaves the GPU idle (the blocking read costs ~50 us of a 1.9 ms
        step: host-side dict building, index copies and the graph launch all sit behind the synchronisation)."""
        vec = torch.cat([w['acc'], w['total']])
        if self.world > 1:
            torch.distributed.all_reduce(vec, group=self.dp_group)
            vec /= self.world
        if not hasattr(self, '_pin'):
            self._pin = [torch.empty(9, dtype=torch.float32).pin_memory() for _ in range(8)]
            self._pin_owner = [None] * 8
            self._pin_i = 0
        i = self._pin_i = (self._pin_i + 1) % 8
        old = self._pin_owner[i]() if self._pin_owner[i] is not None else None
        if old is not None:
            old.result()              # still held eight steps later: take the values out before the slot is rewritten
        self._pin[i].copy_(vec, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        out = PendingLosses(self._pin[i], ev, norm, keys)
        self._pin_owner[i] = weakref.ref(out)
        return out

    @staticmethod
    def loss_dict_from_host(host, norm):
        n = norm
        d = dict(kp_loss=host[4], vp_recon_loss=host[0] * n[0], vp_kl_loss=host[1] * n[1], gmm_loss=host[3] * n[3],
                 loss_3d=host[2], instance_loss=host[5], total_loss=host[8])
        return {k: np.asarray(v, dtype=np.float32) for k, v in d.items()}

    @torch.no_grad()
    @_on_device
    def predict(self, view_idx, frame_idx, add_trans=True, phases=None, want_vertices=True):
        """get_preds_batch forward (:3637-3672): vertices (N,6890,3), the 25 loss joints, poses, orient, trans."""
        C = self.C_
        N = int(view_idx.shape[0])
        w = self.ws(N)
        self._flush_info()
        w['vi'].copy_(view_idx)
        if phases is None:
            w['fi'].copy_(frame_idx)
            torch.index_select(self.phase_table, 0, w['fi'], out=w['raw'])
        else:
            w['raw'].copy_(phases)
        use_noise = self.version >= 3 and self.training and self.args.code_noise > 0 and self.Cc > 0
        self._motion_forward(w, N, w['vi'], w['raw'], use_noise)
        self._shape(w)
        self._joints_forward(w, N)
        ro = self.ro
        V = None
        if want_vertices:
            V = torch.empty(N, NV, 3, device=self.dev)
            pose_blend(w['F_m'], C, w['PO_m'], N, LDF)
            lib.skin_fwd(w['A_m'], w['PO_m'], w['base'], C.Wt, V, N)
        j = torch.empty(N, 25, 3, device=self.dev)
        lib.joints_gather_fwd(w['JP'], w['JX'], ro['np'] + ro['nx'], ro['np'], self.map25, 25,
                              w['trans'] if add_trans else None, None, j, N)
        if want_vertices and add_trans:
            V += w['trans'][:, None]
        return dict(v=V, j=j, poses=w['aa'][:, 3:].clone(), orient=w['Y'][:N, :6].clone(),
                    orient_aa=w['aa'][:, :3].clone(), trans=w['trans'].clone())
