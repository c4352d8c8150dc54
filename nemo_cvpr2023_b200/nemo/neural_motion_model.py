"""Mirror of the fit models of `nemo/neural_motion_model.py`: FCNN (:58), RotNet (:74), MotionNet (:106),
MultiViewModel (:151) and NemoV1..V4 (:3364, :3675, :3786, :3959) with the reference's method names,
arguments, return values and state_dict keys.  `step`, `warmup`, `opt_cam`, `get_preds_batch`,
`get_preds` run on the fused engine (fit_engine.FitEngine): hand-written CUDA forward + backward +
Adam, no autograd graph.  `eval_3d` / `eval_2d` (:1056-1282, :522-710) are batched over the frames of a view (the
reference calls SMPL once per frame) and write the same CSV columns.  Rendering methods are out of scope."""
import os
from collections import defaultdict
from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn

from .. import assets
from .. import functional as Fn
from ..fit_engine import FitEngine
from ..hmr.smpl import SMPL
from ..hmr.smplify.prior import MaxMixturePrior
from ..human_body_prior.models.vposer_model import VPoser
from ..monotonic_network import MonotonicNetwork
from .rbf import RBF
from .utils import GMoF, ravel_first_2dims

MAX_SIZE = 2000


class _ViewLinear(nn.Module):
    """nn.Linear whose parameters are views into the engine's flat buffers."""

    def __init__(self, weight, bias):
        super().__init__()
        self.weight, self.bias = nn.Parameter(weight), nn.Parameter(bias)
        self.in_features, self.out_features = weight.shape[1], weight.shape[0]

    def forward(self, x, act=None):
        return Fn.linear(x, self.weight, self.bias, act)


class FCNN(nn.Module):
    """:58-71.  `net` keeps the Sequential indices 0/2/4 of the reference for state_dict parity."""

    def __init__(self, input_dim=None, h_dim=None, output_dim=None, _views=None, device='cuda'):
        super().__init__()
        if _views is None:
            mk = lambda o, i: (torch.empty(o, i, device=device).uniform_(-1, 1) / i ** 0.5,
                               torch.empty(o, device=device).uniform_(-1, 1) / i ** 0.5)
            _views = [mk(h_dim, input_dim), mk(h_dim, h_dim), mk(output_dim, h_dim)]
        self.net = nn.ModuleDict({'0': _ViewLinear(*_views[0]), '2': _ViewLinear(*_views[1]),
                                  '4': _ViewLinear(*_views[2])})

    def forward(self, x, last_act=None):
        x = self.net['0'](x, 'relu')
        x = self.net['2'](x, 'relu')
        return self.net['4'](x, last_act)


def _head_init(n_joints, fcnn_dim, device, init_last_layer_zero):
    w = torch.empty(n_joints * 6, fcnn_dim, device=device)
    nn.init.xavier_uniform_(w, gain=0.00001 if init_last_layer_zero else 0.01)
    if init_last_layer_zero:
        b = torch.tensor([1, 0, 0, 1, 0, 0], device=device).float().repeat(n_joints)
    else:
        b = torch.empty(n_joints * 6, device=device).uniform_(-1, 1) / fcnn_dim ** 0.5
    return w, b


class RotNet(nn.Module):
    """:74-103 (NemoV0 only)."""

    def __init__(self, input_dim, fcnn_dim, n_joints, init_last_layer_zero=False, device='cuda'):
        super().__init__()
        self.n_joints = n_joints
        self.net = FCNN(input_dim, fcnn_dim, fcnn_dim, device=device)
        self.linear = _ViewLinear(*_head_init(n_joints, fcnn_dim, device, init_last_layer_zero))

    def forward(self, x):
        z = self.net(x, 'relu')
        rot6d = self.linear(z)
        rotmat = Fn.rot6d_to_rotmat(rot6d.reshape(-1, 6)).view(x.shape[0], self.n_joints, 3, 3)
        pose = Fn.rotmat_to_aa(rotmat.reshape(-1, 3, 3), True).reshape(-1, 3 * self.n_joints)
        return {'rot6d': rot6d, 'rotmat': rotmat, 'pose': pose}


class MotionNet(nn.Module):
    """:106-148.  Stand-alone use goes through the autograd ops; inside the fit models the same
    parameters are driven by the fused engine."""

    def __init__(self, input_dim=None, fcnn_dim=None, n_joints=24, n_linear_out=3, init_last_layer_zero=False,
                 _views=None, device='cuda'):
        super().__init__()
        self.n_joints = n_joints
        if _views is None:
            self.net = FCNN(input_dim, fcnn_dim, fcnn_dim, device=device)
            self.rot_out = _ViewLinear(*_head_init(n_joints, fcnn_dim, device, init_last_layer_zero))
            self.linear_out = _ViewLinear(torch.empty(n_linear_out, fcnn_dim, device=device).uniform_(-1, 1) / fcnn_dim ** 0.5,
                                          torch.empty(n_linear_out, device=device).uniform_(-1, 1) / fcnn_dim ** 0.5)
        else:
            g = lambda k: _views['learned_motion.' + k]
            self.net = FCNN(_views=[(g('net.net.%d.weight' % i), g('net.net.%d.bias' % i)) for i in (0, 2, 4)])
            self.rot_out = _ViewLinear(g('rot_out.weight'), g('rot_out.bias'))
            self.linear_out = _ViewLinear(g('linear_out.weight'), g('linear_out.bias'))

    def forward(self, x):
        B = x.shape[0]
        z = self.net(x, 'relu')
        rot6d = self.rot_out(z)
        rotmat = Fn.rot6d_to_rotmat(rot6d.reshape(-1, 6)).view(B, self.n_joints, 3, 3)
        pose = Fn.rotmat_to_aa(rotmat.reshape(-1, 3, 3), True).reshape(-1, 3 * self.n_joints)
        trans = self.linear_out(z)
        orient_dict = {'rot6d': rot6d[:, :6], 'rotmat': rotmat[:, :1], 'pose': pose[:, :3]}
        pose_dict = {'rot6d': rot6d[:, 6:], 'rotmat': rotmat[:, 1:], 'pose': pose[:, 3:]}
        return pose_dict, orient_dict, trans


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam/AdamW-compatible handle (param_groups, state_dict, zero_grad, step) on one flat engine
    buffer.  `state_dict()` / `load_state_dict()` speak torch.optim.Adam's own format (per-parameter `step`,
    `exp_avg`, `exp_avg_sq`, in the reference's parameter order), so checkpoints move between the reference and this
    mirror in both directions (`MultiViewModel.save/load`, nemo/neural_motion_model.py:257-280)."""

    def __init__(self, params, group, lr, weight_decay=0.0, decoupled=False):
        self._g = group
        super().__init__(list(params), dict(lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay,
                                            decoupled=decoupled))

    def zero_grad(self, set_to_none=False):
        self._g.g.zero_()

    def sync_hyper(self):
        pg = self.param_groups[0]
        self._g.lr, self._g.wd, self._g.betas, self._g.eps = pg['lr'], pg['weight_decay'], pg['betas'], pg['eps']

    @torch.no_grad()
    def step(self, closure=None, skip_lag=False):
        self.sync_hyper()
        self._g.adam_step(skip_lag=skip_lag)

    # ---- the engine's storage of optimiser parameter i: (moment views, step counter)
    def _slot(self, p):
        g = self._g
        off = (p.data_ptr() - g.p.data_ptr()) // 4
        if not (0 <= off < g.p.numel()):
            raise RuntimeError('optimizer parameter is not a view of the engine buffer')
        # as_strided takes an offset into the underlying STORAGE; the group's buffers are slices of the engine's
        m = torch.as_strided(g.m, p.shape, p.stride(), g.m.storage_offset() + off)
        v = torch.as_strided(g.v, p.shape, p.stride(), g.v.storage_offset() + off)
        lag = any(lo <= off < lo + k for lo, k, lagging in g.ranges() if lagging)
        return m, v, (g.step_lag_dev if lag else g.step_dev)

    def state_dict(self):
        sd = super().state_dict()
        state = {}
        for i, p in enumerate(self.param_groups[0]['params']):
            m, v, step = self._slot(p)
            n = int(step.item())
            if n > 0:   # torch.optim.Adam creates the state of a parameter at its first step with a gradient
                state[i] = {'step': torch.tensor(float(n)), 'exp_avg': m.clone(), 'exp_avg_sq': v.clone()}
        sd['state'] = state
        sd['param_groups'][0].pop('decoupled', None)
        return sd

    def load_state_dict(self, sd):
        g = self._g
        if 'flat' in sd:   # round-1 format of this repo
            g.m.copy_(sd['flat']['m']), g.v.copy_(sd['flat']['v']), g.step_dev.copy_(sd['flat']['step'])
            if hasattr(g, 'step_lag_dev'):
                g.step_lag_dev.copy_(sd['flat']['step'])
        else:
            g.reset_state()
            steps = {}
            for i, p in enumerate(self.param_groups[0]['params']):
                st = sd['state'].get(i)
                if st is None:
                    continue
                m, v, step = self._slot(p)
                m.copy_(torch.as_tensor(st['exp_avg']).reshape(p.shape))
                v.copy_(torch.as_tensor(st['exp_avg_sq']).reshape(p.shape))
                n = int(st['step'])
                if steps.setdefault(step.data_ptr(), n) != n:
                    raise RuntimeError('checkpoint has per-parameter Adam step counts this engine cannot represent '
                                       '(%d vs %d within one step class)' % (steps[step.data_ptr()], n))
                step.fill_(n)
        if sd.get('param_groups'):
            for k in ('lr', 'weight_decay', 'betas', 'eps'):
                if k in sd['param_groups'][0]:
                    self.param_groups[0][k] = sd['param_groups'][0][k]
        self.sync_hyper()


class _StepInfo(dict):
    """`info_dict` of `step` (view_idx, frame_idx, loss_all (N,25,2), points2d_gt (N,25,3)).  The tensors live in the
    engine's reused workspace; an entry is copied out of it when it is first read (the fit loop of the reference reads
    them only when it logs, scripts/learned_multi_view_recon_nn.py:247-308), which keeps four copy kernels out of an
    iteration whose info_dict is dropped.  An info_dict that is still referenced when the model steps again is filled
    in just before the workspace is overwritten (`FitEngine._flush_info`), so it never shows a later step's data."""
    _LAZY = ('view_idx', 'frame_idx', 'loss_all', 'points2d_gt')

    def __init__(self, engine, w, gt_all):
        super().__init__()
        self._e, self._w, self._gt, self._tick = engine, w, gt_all, engine.step_tick
        engine.track_info(self)

    def materialize(self):
        for k in self._LAZY:
            self[k]

    def __missing__(self, k):
        if k not in self._LAZY:
            raise KeyError(k)
        if self._e.step_tick != self._tick:
            raise RuntimeError('info_dict[%r] of an earlier step was first read after the model stepped again' % k)
        w = self._w
        v = {'view_idx': lambda: w['vi'].clone(), 'frame_idx': lambda: w['fi'].clone(),
             'loss_all': lambda: w['loss_all'].clone(), 'points2d_gt': lambda: self._gt[w['vi'], w['fi']]}[k]()
        self[k] = v
        return v

    def __contains__(self, k):
        return k in self._LAZY or dict.__contains__(self, k)

    def keys(self):
        return list(self._LAZY)

    def __iter__(self):
        return iter(self._LAZY)

    def __len__(self):
        return len(self._LAZY)

    def items(self):
        return [(k, self[k]) for k in self._LAZY]

    def values(self):
        return [self[k] for k in self._LAZY]

    def get(self, k, default=None):
        return self[k] if k in self._LAZY else default


class MultiViewModel(nn.Module):
    """:151-255 + the fit-path methods (:2758-3124).  Extra keyword arguments replace the licensed files the
    reference loads from `software/`: `smpl_arrays`, `vposer_state`, `gmm` (default: seeded synthetic
    stand-ins of the exact shapes, see assets.py)."""

    VERSION = 1

    def __init__(self, args, multi_view_seqs, device='cuda', smpl_arrays=None, vposer_state=None, gmm=None,
                 init_params=None):
        super().__init__()
        self.args = args
        if not hasattr(self.args, 'include_vs'):
            self.args.include_vs = False
            self.args.include_pare = False
        self.FOCAL_LENGTH = assets.FOCAL_LENGTH
        self.IMG_D0, self.IMG_D1 = multi_view_seqs.IMG_D0, multi_view_seqs.IMG_D1
        self.n_joints = 23
        self.device = device
        self.multi_view_seqs = multi_view_seqs
        self.num_views, self.num_frames = multi_view_seqs.num_views, multi_view_seqs.num_frames
        self.smpl = SMPL(smpl_arrays if smpl_arrays is not None else 'synthetic', batch_size=1, create_transl=False,
                         device=device)
        self.criterion_keypoints = nn.MSELoss(reduction='none')
        self.robustifier = GMoF()
        vposer_state = vposer_state if vposer_state is not None else assets.make_vposer_weights(2)
        self.vp = VPoser(state_dict=vposer_state, device=device)
        self.l1_loss = nn.L1Loss(reduction='mean')
        gmm = gmm if gmm is not None else assets.make_gmm(3)
        self.pose_prior = MaxMixturePrior(num_gaussians=8, gmm=gmm, device=device)
        self.points2d_gt_all, self.gt_bbox_size = self.collate_gt_2d()
        hmr = np.stack([np.asarray(multi_view_seqs.sequences[v]['pose']) for v in range(self.num_views)])
        self.hmr_theta = torch.tensor(hmr[:, :, 3:-1], dtype=torch.float32, device=device)
        self.hmr_mask = torch.tensor(hmr[:, :, -1:], dtype=torch.float32, device=device)
        data = dict(points2d_gt_all=self.points2d_gt_all.cpu().numpy(), gt_bbox_size=self.gt_bbox_size.cpu().numpy(),
                    hmr_theta=hmr[:, :, 3:-1], hmr_mask=hmr[:, :, -1:], IMG_D0=self.IMG_D0, IMG_D1=self.IMG_D1)
        if init_params is None:
            init_params = self._reference_init()
        self.engine = FitEngine(self.args, self.num_views, self.num_frames, self.smpl.consts, vposer_state, gmm, data,
                                version=self.VERSION, device=device, init_params=init_params)
        self._bind_parameters()
        self.training = False

    # ---- construction helpers
    def _reference_init(self):
        """Initial values as NemoV1/V2.__init__ draws them (:3375-3402, :3685-3692)."""
        a = self.args
        rbf_dim = a.phase_rbf_dim if self.VERSION >= 2 else 0
        shadow = deepcopy(a)
        shadow.phase_rbf_dim = rbf_dim
        p = assets.make_fit_params(shadow, self.num_views, seed=int(torch.randint(0, 2 ** 31 - 1, (1,))),
                                   img_d0=self.IMG_D0, rot_out_std=0.0)
        w = torch.empty(144, a.h_dim)
        nn.init.xavier_uniform_(w, gain=0.00001)
        p['learned_motion.rot_out.weight'] = w.numpy()
        return p

    def _bind_parameters(self):
        e = self.engine
        P = nn.Parameter
        self.learned_cameras = P(e.g_cam.views['learned_cameras'])
        if e.g_inst is not None:
            self.learned_instance_code = P(e.g_inst.views['learned_instance_code'])
        self.learned_motion = MotionNet(_views=e.g_motion.views)
        self.learned_betas = P(e.betas)
        nets = []
        for v in range(self.num_views):
            m = MonotonicNetwork.__new__(MonotonicNetwork)
            nn.Module.__init__(m)
            m.n_nodes = e.K
            m.shifts, m.scales = P(e.g_phase.views['shifts'][v]), P(e.g_phase.views['scales'][v])
            nets.append(m)
        self.phase_networks = nn.ModuleList(nets)
        if e.D > 0:
            rbf = RBF.__new__(RBF)
            nn.Module.__init__(rbf)
            rbf.in_features, rbf.out_features, rbf.basis_name = 1, e.D, self.args.rbf_kernel
            rbf.register_buffer('centres', e.centres.view(-1, 1))
            rbf.log_sigmas = P(e.g_motion.views['phase_rbf.log_sigmas'])
            self.phase_rbf = rbf
        # .grad of every parameter aliases the engine's flat gradient buffers
        grads = e.named_grads()
        for k, p in self.named_parameters():
            if k in grads:
                p.grad = grads[k]
        a = self.args
        self.opt_cameras = FusedAdam([self.learned_cameras], e.g_cam, a.lr_camera)
        motion = list(self.learned_motion.parameters()) + (list(self.phase_rbf.parameters()) if e.D > 0 else [])
        self.opt_motion = FusedAdam(motion, e.g_motion, a.lr_human, a.wd_human, decoupled=(a.opt_human == 'adamw'))
        self.opt_phase = FusedAdam(self.phase_networks.parameters(), e.g_phase, a.lr_phase)
        self.optimizers = [self.opt_cameras, self.opt_motion, self.opt_phase]
        if e.g_inst is not None:
            self.opt_instance = FusedAdam([self.learned_instance_code], e.g_inst, a.lr_instance)
            self.optimizers.append(self.opt_instance)
        if a.lr_factor < 1:
            self.schedulers = [torch.optim.lr_scheduler.ReduceLROnPlateau(o, factor=a.lr_factor, min_lr=1e-6)
                               for o in self.optimizers]

    def _apply(self, fn, recurse=True):
        # parameters are views into the engine's flat CUDA buffers: moving the module would detach them
        return self

    # ---- data
    def collate_gt_2d(self, label_type=None):
        """:2908-2961"""
        label_type = label_type or self.args.label_type
        key = {'op': 'pose_2d_op', 'gt': 'pose_2d_gt', 'vibe': 'vibe_joints2d', 'pare': 'pare_joints2d',
               'vs': 'vs_joints2d'}
        gt = []
        for v in range(self.num_views):
            seq = self.multi_view_seqs.sequences[v]
            if label_type == 'intersection':
                g1, g2 = np.array(seq['pose_2d_op']), np.array(seq['pose_2d_gt'])
                mean = (g1 + g2)[..., :2] / 2
                dist = np.sqrt(np.power(g1[..., :2] - g2[..., :2], 2).sum(-1, keepdims=True))
                conf = (dist < self.args.label_intersection_threshold).astype('float32') * g1[..., -1:]
                gt.append(np.concatenate([mean, conf], -1))
            else:
                gt.append(seq[key[label_type]])
        pts = torch.tensor(np.array(gt)).float().to(self.device)
        d0 = pts[..., 0].max(-1)[0] - pts[..., 0].min(-1)[0]
        d1 = pts[..., 1].max(-1)[0] - pts[..., 1].min(-1)[0]
        return pts, torch.sqrt(d0 ** 2 + d1 ** 2) + 1e-4

    # ---- checkpointing (:257-280)
    def save(self, path):
        torch.save({'model_sd': self.state_dict(), 'opt_sd': [o.state_dict() for o in self.optimizers]}, path)

    def load(self, path):
        saved = torch.load(path, weights_only=False)
        sd = {k: v for k, v in saved['model_sd'].items()
              if not any(k.startswith(m) for m in ('vp', 'pose_prior', 'renderer', 'smpl'))}
        self.load_state_dict(sd, strict=False)
        for i, opt in enumerate(self.optimizers):
            opt.load_state_dict(saved['opt_sd'][i])

    # ---- forward API
    def frame_idx_to_raw_phase(self, frame_idx):
        """:2978-2984"""
        return self.engine.phase_table[frame_idx.to(self.engine.dev)]

    def _idx(self, t):
        return torch.as_tensor(t).long()

    def get_preds_batch(self, view_idx, frame_idx, add_trans=True, phases=None, detach_pose=False):
        """:3637-3672 / :3968-4008.  Tensors are results of the fused forward (no autograd graph)."""
        self.engine.training = self.training
        out = self.engine.predict(self._idx(view_idx), None if frame_idx is None else self._idx(frame_idx),
                                  add_trans=add_trans, phases=phases)
        out['view_idx'], out['frame_idx'] = view_idx, frame_idx
        return out

    def get_preds_given_phases_and_view(self, view_idx, input_phases, add_trans=True, phases=None):
        """:3733-3781 / :3911-3956 / :4010-4058: the motion field evaluated at ALREADY WARPED phases `input_phases` (N,1)
        for the instances `view_idx` -> (vertices (N,6890,3), joints (N,49,3), poses (N,69), orient 6D (N,6),
        orient axis-angle (N,3), trans (N,3)).  Runs on the autograd ops (RBF, Linear, rot6d, rotmat->aa, SMPL), so
        like the reference's it can be differentiated; the fit itself goes through the fused engine."""
        a, dev = self.args, self.engine.dev
        view_idx = self._idx(view_idx).to(dev)
        x = torch.as_tensor(input_phases, dtype=torch.float32, device=dev).reshape(-1, 1)
        n = x.shape[0]
        has_rbf = self.VERSION >= 2 and a.phase_rbf_dim > 0
        feats = self.phase_rbf(x) if has_rbf else x
        if a.instance_code_size > 0:
            code = self.learned_instance_code[view_idx]
            if self.VERSION >= 3 and self.training and a.code_noise > 0:   # :3921-3923
                code = code + a.code_noise * torch.randn_like(code)
            feats = torch.cat([feats, code], 1)
        pose_dict, orient_dict, trans = self.learned_motion(feats.contiguous())
        out = self._get_smpl_given_poses(pose_dict['rotmat'].contiguous(), orient_dict['rot6d'].contiguous(), 'rotmat')
        zero = torch.zeros(1, 1, device=dev)
        inp0 = torch.cat([self.phase_rbf(zero) if has_rbf else zero,
                          torch.zeros(1, a.instance_code_size, device=dev)], 1)
        trans_0 = self.learned_motion(inp0.contiguous())[-1]
        if not getattr(self, 'start_global_traj_anywhere', False):
            trans = trans - trans_0
        v, j = out.vertices, out.joints
        if add_trans:
            if getattr(self, 'detach_articulation', False):
                v, j = v.detach(), j.detach()
            v, j = v + trans.unsqueeze(1), j + trans.unsqueeze(1)
        return v, j, pose_dict['pose'], orient_dict['rot6d'], orient_dict['pose'], trans

    def get_preds(self, add_trans=True):
        """:2986-3003: one get_preds_batch per view, stacked to (V, F, ...)."""
        all_preds = defaultdict(list)
        for i in range(self.num_views):
            vi = i * torch.ones(self.num_frames).long().to(self.device)
            fi = torch.arange(self.num_frames).long().to(self.device)
            for k, v in self.get_preds_batch(vi, fi, add_trans=add_trans).items():
                all_preds[k].append(v)
        return {k: torch.stack(v) for k, v in all_preds.items()}

    # ---- evaluation: same numbers and CSV columns as the reference, one SMPL call per view
    def _eval_frames(self, num_frames, num_views, view_idxs):
        ncol = self.num_frames if num_frames < 0 else min(self.num_frames, num_frames)
        if len(view_idxs) == 0:
            view_idxs = np.arange(0, self.num_views if num_views < 0 else min(self.num_views, num_views))
        frames = [int(np.round(c / ncol * self.num_frames)) for c in range(ncol)]
        return list(view_idxs), frames

    def _smpl_of_body_pose(self, body_pose):
        """SMPL(betas=None, body_pose=aa (B,69), global_orient=None, pose2rot=True) as eval_3d calls it (:1144-1151)."""
        bp = torch.as_tensor(np.asarray(body_pose) if not isinstance(body_pose, torch.Tensor) else body_pose,
                             dtype=torch.float32).to(self.device).reshape(-1, 69)
        with torch.no_grad():  # the zero global orientation / betas of the reference call, at the batch size
            out = self.smpl(betas=None, body_pose=bp, global_orient=torch.zeros(bp.shape[0], 3, device=bp.device),
                            pose2rot=True)
        return out.vertices.contiguous(), out.joints[:, :15].contiguous()

    def eval_3d(self, out_dir, num_frames=-1, num_views=-1, view_idxs=[], dynamic_only=False):
        """:1056-1282.  MPVPE / MPJPE (mm, no Procrustes) of the fitted poses -- and of every baseline pose track the
        sequences carry (`pose` = VIBE, `vs_pose`, `pare_pose`, `glamr_pose`) -- against `pose_3d_gt`."""
        import pandas as pd
        from .utils.pose_utils import reconstruction_error
        view_idxs, frames = self._eval_frames(num_frames, num_views, view_idxs)
        pred_dict = self.get_preds()
        os.makedirs(out_dir, exist_ok=True)
        stats = defaultdict(list)
        for view_idx in view_idxs:
            seq = self.multi_view_seqs.sequences[view_idx]
            gt_pose = np.asarray(torch.stack(list(seq['pose_3d_gt'])).cpu() if isinstance(seq['pose_3d_gt'], (list, tuple))
                                 else seq['pose_3d_gt'], dtype=np.float32)
            fsel = list(frames)
            if dynamic_only:  # :1082-1116: frames between the first and the last fast GT joint motion
                _, j_all = self._smpl_of_body_pose(gt_pose[:, 3:])
                j_all = j_all.cpu().numpy()
                vel = np.sqrt(((j_all[1:] - j_all[:-1]) ** 2).sum(-1)) * (
                    30 * self.multi_view_seqs.framerate_multiplier[view_idx])
                inds = np.where(vel.max(1) >= 2)[0]
                mask = np.zeros((self.num_frames,))
                if len(inds):
                    mask[inds.min():inds.max()] = 1
                fsel = [f for f in frames if mask[f] != 0]
            if not fsel:
                continue
            v_gt, j_gt = self._smpl_of_body_pose(gt_pose[fsel][:, 3:])
            tracks = {'ours': pred_dict['poses'][view_idx][fsel]}
            if 'pose' in seq:
                tracks['vibe'] = np.asarray(seq['pose'], np.float32)[fsel][:, 3:-1]
            for name, key, sl in (('vs', 'vs_pose', slice(3, -1)), ('pare', 'pare_pose', slice(3, -1)),
                                  ('glamr', 'glamr_pose', slice(None, -1))):
                if key in seq:
                    tracks[name] = np.asarray(seq[key], np.float32)[fsel][:, sl]
            for name, bp in tracks.items():
                v, j = self._smpl_of_body_pose(bp)
                stats['mpjpe-' + name].append(1000 * reconstruction_error(j_gt, j, pa=False))
                stats['mpvpe-' + name].append(1000 * reconstruction_error(v_gt, v, pa=False))
            print(f"View: {view_idx}, MPVPE: {stats['mpvpe-ours'][-1]:.2f}, MPJPE: {stats['mpjpe-ours'][-1]:.2f}")
        df = pd.DataFrame(stats)
        df.to_csv(os.path.join(out_dir, 'eval_3d.csv' if not dynamic_only else 'eval_3d_dynamic.csv'))
        return df

    def eval_2d(self, out_dir, num_frames=-1, num_views=-1, view_idxs=[]):
        """:522-710.  ReconError2D (`rmse` keypoint loss) and PCK@0.05 of the re-projected fitted joints -- and of every
        2-D baseline track the sequences carry -- against the `gt` 2-D labels (first 15 joints)."""
        import pandas as pd
        view_idxs, frames = self._eval_frames(num_frames, num_views, view_idxs)
        gt_all, bbox = self.collate_gt_2d(label_type='gt')
        base = {'op': self.collate_gt_2d(label_type='op')[0]}
        seq0 = self.multi_view_seqs.sequences[0]
        for name, key, on in (('vibe', 'vibe_joints2d', True), ('vs', 'vs_joints2d', self.args.include_vs),
                              ('pare', 'pare_joints2d', self.args.include_pare)):
            if on and key in seq0:
                base[name] = self.collate_gt_2d(label_type=name)[0]
        pred = self.get_preds()
        V, F = pred['j'].shape[:2]
        with torch.no_grad():
            p2d = self.learned_camera_projection(pred['j'].reshape(-1, 25, 3), pred['view_idx'].reshape(-1))
        base = dict(ours=p2d.reshape(V, F, 25, 2), **base)
        os.makedirs(out_dir, exist_ok=True)
        stats = defaultdict(list)
        fi = torch.as_tensor(frames, device=gt_all.device)
        for view_idx in view_idxs:
            gt = gt_all[view_idx][fi][:, :15].contiguous()
            size = bbox[view_idx][fi].contiguous()
            for name, pts in base.items():
                p = pts[view_idx][fi][:, :15, :2].float().contiguous()
                r = self.keypoint_loss(p, gt[..., :2], gt[..., 2:], loss_type='rmse').mean().item()
                hits = torch.zeros(2, device=p.device)
                Fn.lib.pck2d(p, gt, size, 15, p.shape[0], hits)
                stats['recon_error_2d-' + name].append(r)
                stats['pck-' + name].append(float(100 * hits[0] / hits[1]))
            print(f"View: {view_idx}, ReconError2D: {stats['recon_error_2d-ours'][-1]:.2f}, PCK: {stats['pck-ours'][-1]:.1f}")
        df = pd.DataFrame(stats)
        df.to_csv(os.path.join(out_dir, 'eval_2d.csv'))
        return df

    def _get_smpl_given_poses(self, pose_input, orient, pose_type):
        """:2963-2976"""
        if pose_type == 'aa':
            rot = Fn.batch_rodrigues(pose_input.reshape(-1, 3), kind=0).reshape(-1, 23, 3, 3)
        elif pose_type == 'rotmat':
            rot = pose_input
        return self.smpl(betas=self.learned_betas, body_pose=rot,
                         global_orient=Fn.rot6d_to_rotmat(orient).unsqueeze(1), pose2rot=False)

    def learned_camera_projection(self, input_points3d, view_idx):
        """:3073-3124, batched: the camera of each sample is looked up by view_idx."""
        cam = self.learned_cameras[view_idx.to(self.learned_cameras.device).long()]
        rot = Fn.rot6d_to_rotmat(cam[:, 3:].contiguous())
        n = input_points3d.shape[0]
        dev = input_points3d.device
        centre = torch.tensor([[float(self.IMG_D0 // 2), float(self.IMG_D1 // 2)]], device=dev).expand(n, 2)
        return Fn.perspective_projection(input_points3d, rot, cam[:, :3].contiguous(),
                                         torch.full((n,), self.FOCAL_LENGTH, device=dev), centre, 0.0)

    def keypoint_loss(self, pred_keypoints_2d, gt_keypoints_2d, gt_weight, gt_size=None, loss_type=None):
        """:2806-2843 (API form; inside step() the same arithmetic is fused into the projection kernel)."""
        loss_type = loss_type or self.args.loss
        m = (gt_weight > 0.5).float()
        if loss_type in ('rmse', 'rmse_resized'):
            if loss_type == 'rmse_resized':
                s = gt_size.unsqueeze(-1).unsqueeze(-1)
                pred_keypoints_2d, gt_keypoints_2d = pred_keypoints_2d / s, gt_keypoints_2d / s
            return m * torch.sqrt(1e-6 + self.criterion_keypoints(pred_keypoints_2d, gt_keypoints_2d).sum(-1, keepdim=True))
        if loss_type == 'mse':
            return m * self.criterion_keypoints(pred_keypoints_2d, gt_keypoints_2d)
        if loss_type == 'rmse_robust':
            return m * self.robustifier(pred_keypoints_2d - gt_keypoints_2d, sqrt=True)
        if loss_type == 'mse_robust':
            return m * self.robustifier(pred_keypoints_2d - gt_keypoints_2d, sqrt=False)
        if loss_type == 'mse_robust_resized':
            s = gt_size.unsqueeze(-1).unsqueeze(-1)
            return m * self.robustifier(pred_keypoints_2d / s * 1000 - gt_keypoints_2d / s * 1000, sqrt=False)
        raise ValueError(loss_type)

    def gmm_prior_loss(self, poses, orient_aa, betas):
        """:2758-2773"""
        return self.pose_prior(poses, betas).mean()

    def vposer_loss(self, poses, orient):
        """:2775-2804 (API form on the autograd ops)."""
        N = poses.size(0)
        q_z = self.vp.encode(poses[:, :63].contiguous())
        dec = self.vp.decode(q_z.mean)['pose_body'].contiguous().view(N, -1)
        recon = torch.cat([dec, poses[:, 63:]], 1)
        bm_orig = self._get_smpl_given_poses(poses, orient, 'aa')
        bm_rec = self._get_smpl_given_poses(recon.contiguous().view(N, -1), orient, 'aa')
        v2v = self.l1_loss(bm_rec.vertices.detach(), bm_orig.vertices)
        p_z = torch.distributions.normal.Normal(torch.zeros_like(q_z.mean), torch.ones_like(q_z.mean))
        kl = torch.mean(torch.sum(torch.distributions.kl.kl_divergence(q_z, p_z), dim=[1]))
        return v2v, kl

    # ---- optimisation
    _LOSS_KEYS = ['kp_loss', 'gmm_loss', 'vp_recon_loss', 'vp_kl_loss', 'total_loss']

    def _full_indices(self):
        v = torch.arange(self.num_views).repeat_interleave(self.num_frames)
        f = torch.arange(self.num_frames).repeat(self.num_views)
        return v, f

    def step(self, view_idx, frame_idx, update=True, full_batch=False, sync=True):
        """:3511-3598 / :3796-3909 -> (loss_dict of numpy scalars, info_dict).
        sync=False (not in the reference): the losses come back as a `PendingLosses` -- their device->host copy is queued
        into pinned memory and `result()` gives the same dict once it has landed, so a loop that logs the losses of step
        i while step i + 1 runs keeps the GPU busy.  A plateau scheduler (lr_factor < 1) needs the value at once and
        waits for it either way."""
        e = self.engine
        if update and self.VERSION >= 3:
            self.training = True
        e.training = self.training
        if not (self.args.batch_size > -1 and not full_batch):
            view_idx, frame_idx = self._full_indices()
        vi, fi = self._idx(view_idx), self._idx(frame_idx)
        fused = update and all(isinstance(o, FusedAdam) for o in self.optimizers)
        if fused:      # forward + backward + Adam as one CUDA-graph launch (eager under data parallelism)
            for o in self.optimizers:
                o.sync_hyper()
            w, norm = e.step_graph(vi, fi, mode='step', which=[o._g for o in self.optimizers])
        else:
            w, norm = e.forward_backward(vi, fi, update=update, mode='step')
            if update:
                self._step_optimizers()
        if sync:
            ld = e.loss_dict(w, norm)
            loss_dict = {k: ld[k] for k in self._LOSS_KEYS}
        else:
            loss_dict = e.loss_dict_async(w, norm, self._LOSS_KEYS)
        if update and self.args.lr_factor < 1:
            total = ld['total_loss'] if sync else loss_dict.result()['total_loss']
            for s in self.schedulers:
                s.step(float(total))
        info_dict = _StepInfo(e, w, self.points2d_gt_all)
        self.training = False
        e.training = False
        return loss_dict, info_dict

    def _step_optimizers(self, which=None, skip_lag=False):
        """`for opt in self.optimizers: opt.step()` (:3586-3588) as one launch over every flat buffer."""
        opts = which if which is not None else self.optimizers
        if all(isinstance(o, FusedAdam) for o in opts):
            for o in opts:
                o.sync_hyper()
            self.engine.optimizer_step([o._g for o in opts], skip_lag=skip_lag)
        else:
            for o in opts:
                o.step()

    def _sample(self):
        bs = self.args.batch_size
        return torch.randint(0, self.num_views, size=(bs,)), torch.randint(0, self.num_frames, size=(bs,))

    def warmup(self, warmup_steps=1000):
        """:3455-3509: fit the predicted axis-angle poses to the VIBE thetas (GMoF), stepping the motion
        and phase optimisers only.  (The reference also runs the unused LBS here; it is skipped.)"""
        if warmup_steps == 0:
            return []
        if not self.args.batch_size > -1:
            raise NotImplementedError()
        e = self.engine
        e.training = self.training
        losses = []
        for _ in range(warmup_steps):
            vi, fi = self._sample()
            # linear_out.* get no gradient here: torch.optim.Adam skips them (no weight decay, no state, step stays 0)
            for o in (self.opt_motion, self.opt_phase):
                o.sync_hyper()
            w, _ = e.step_graph(vi, fi, mode='warmup', which=[self.opt_motion._g, self.opt_phase._g], skip_lag=True)
            losses.append(w['acc'][2].clone())   # `acc` is a reused workspace slot
        out = [float(x) for x in torch.stack(losses).cpu()]
        # the reference stops on the first NaN gradient or collapsed pose (:3497-3508); here the loop never leaves the
        # device, so the check happens once, on the loss trace and on the parameters the warm-up stepped
        if not (np.isfinite(out).all() and bool(torch.isfinite(e.g_motion.p).all()) and bool(torch.isfinite(e.g_phase.p).all())):
            raise FloatingPointError('nan gradient found during warm-up (reference: nemo/neural_motion_model.py:3497-3500)')
        return out

    def opt_cam(self, cam_opt_steps=2000):
        """:2869-2906: Adam on the cameras only, first frame of every view, plain-mean reprojection loss."""
        e = self.engine
        vi, fi = torch.arange(self.num_views), torch.zeros(self.num_views, dtype=torch.long)
        e.training = self.training
        # the reference builds a throw-away Adam here (:2873-2876): self.opt_cameras starts fresh afterwards, so the
        # camera fit runs on an alias of the camera buffers with its own moments and step count
        tmp = e.g_cam.fresh_alias()
        tmp.lr = self.args.lr_camera
        log = []
        try:
            for _ in range(cam_opt_steps):
                w, _ = e.step_graph(vi, fi, mode='cam_fit', which=[tmp])
                log.append(w['acc'][4].clone())
        finally:
            e.drop_graphs(tmp)
        return [np.asarray(x) for x in torch.stack(log).cpu().numpy()] if log else []


class NemoV1(MultiViewModel):
    """:3364-3672: a single network for pose + orient + trans."""
    VERSION = 1


class NemoV2(NemoV1):
    """:3675-3781: V1 + RBF phase features."""
    VERSION = 2

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.detach_articulation = False
        self.start_global_traj_anywhere = False


class NemoV3(NemoV2):
    """:3786-3956: V2 + instance-code regulariser + 3D pose loss (+ optional code noise while training)."""
    VERSION = 3
    _LOSS_KEYS = ['kp_loss', 'gmm_loss', 'vp_recon_loss', 'instance_loss', 'loss_3d', 'vp_kl_loss', 'total_loss']


class NemoV4(NemoV3):
    """:3959-4149: V3 with OP-Nose as joint 0 and the batch opt_cam that steps every optimiser."""
    VERSION = 4

    def opt_cam(self, cam_opt_steps=2000):
        if not self.args.batch_size > -1:
            raise NotImplementedError()
        e = self.engine
        e.training = self.training
        for _ in range(cam_opt_steps):
            vi, fi = self._sample()
            for o in self.optimizers:
                o.sync_hyper()
            e.step_graph(vi, fi, mode='opt_cam', detach_pose=True, which=[o._g for o in self.optimizers])


class NemoV0(MultiViewModel):
    """:3127-3361 (separate RotNet / RotNet / FCNN for pose, orientation and translation).  Dead code in the reference:
    the class defines no `get_preds_batch` (its `step` :3277 and `warmup` :3243 call it), `step` adds
    `weight_vp_loss * vposer_loss(...)` where `vposer_loss` returns a tuple (:3337, a TypeError), `warmup` reads a
    `spin_theta` track no loader provides, and no shipped script or config selects it.  There is no behaviour to
    reproduce, so it is not part of the accelerated path; `RotNet` / `FCNN` themselves are available above."""

    def __init__(self, *a, **k):
        raise NotImplementedError('NemoV0 cannot run in the reference either (see the class docstring); use NemoV1..V4')
