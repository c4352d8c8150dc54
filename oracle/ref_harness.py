"""Import and drive the UNMODIFIED reference (/root/reference) in the build container.

Only `oracle/make_golden.py` uses this (to produce `tests/golden/*.npz`) and it only
works where /root/reference is mounted; nothing on the GPU box imports it.
Recipe = SURVEY.md section 8(c):
  1. MagicMock stubs for the visualisation / debugger modules the reference imports at
     module top (`nemo/neural_motion_model.py:1-53`);
  2. a ~30-line `smplx` shim (smplx 0.1.28 is neither vendored nor installed): SMPL.forward
     restated from the smplx API, calling the REFERENCE's vendored lbs
     (`human_body_prior/body_model/lbs.py`);
  3. a temp cwd holding a synthetic `software/spin_data/J_regressor_extra.npy`
     (read by `hmr/smpl.py:23`);
  4. the fit model built with `cls.__new__` + hand-set attributes, bypassing
     `MultiViewModel.__init__` (needs licensed files, hard-codes `.to('cuda')`);
  5. the reference's own `step`, `get_preds_batch`, `vposer_loss`, ... are then called as is.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from collections import namedtuple
from unittest.mock import MagicMock

import numpy as np
import torch
import torch.nn as nn

# the mounted reference in the build container; on the GPU box the byte-identical copy of the path's files that
# `python -m oracle.build_ref` made (git-ignored, shipped with the snapshot)
_LOCAL = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
REF = '/root/reference' if os.path.isdir('/root/reference/nemo') else _LOCAL
_state = {}


def available():
    return os.path.isdir(os.path.join(REF, 'nemo'))


def _install_stubs():
    for name in ['matplotlib', 'matplotlib.pylab', 'matplotlib.pyplot', 'matplotlib.colors',
                 'matplotlib.cm', 'matplotlib.patches', 'matplotlib.animation',
                 'ipdb', 'pyrender', 'pyrender.constants', 'trimesh', 'dotmap',
                 'mpl_toolkits', 'mpl_toolkits.mplot3d', 'skimage', 'skimage.io', 'skimage.util',
                 'skimage.util.shape', 'skimage.transform', 'yacs', 'yacs.config', 'torchgeometry',
                 'torchvision', 'torchvision.transforms', 'torchvision.models',
                 'torchvision.models.resnet', 'torchvision.transforms.functional', 'PIL', 'PIL.Image',
                 'torchvision.utils', 'h5py', 'pytorch3d', 'chumpy', 'smplx.utils']:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock(name=name)


def _install_smplx_shim(sm_np):
    """smplx.SMPL restated (smplx 0.1.28 body_models.py SMPL.forward + VertexJointSelector)
    on top of the reference's vendored lbs."""
    from human_body_prior.body_model import lbs as ref_lbs
    out_t = namedtuple('ModelOutput', ['vertices', 'joints', 'full_pose', 'betas', 'global_orient',
                                       'body_pose'])

    class SMPL(nn.Module):
        def __init__(self, model_path=None, batch_size=1, create_transl=False, **kw):
            super().__init__()
            for k in ['v_template', 'shapedirs', 'posedirs', 'J_regressor', 'lbs_weights']:
                self.register_buffer(k, torch.tensor(sm_np[k]))
            self.register_buffer('parents', torch.tensor(sm_np['parents']).long())
            self.register_buffer('extra_joints_idxs', torch.tensor(sm_np['vertex_joint_ids']).long())
            self.faces = sm_np['faces']
            self.batch_size = batch_size

        def forward(self, betas=None, body_pose=None, global_orient=None, transl=None,
                    return_verts=True, return_full_pose=False, pose2rot=True, **kwargs):
            dt, dev = self.v_template.dtype, self.v_template.device
            if global_orient is None:
                global_orient = torch.zeros(self.batch_size, 3, dtype=dt, device=dev)
            if body_pose is None:
                body_pose = torch.zeros(self.batch_size, 69, dtype=dt, device=dev)
            if betas is None:
                betas = torch.zeros(self.batch_size, 10, dtype=dt, device=dev)
            full_pose = torch.cat([global_orient, body_pose], dim=1)
            b = max(betas.shape[0], global_orient.shape[0], body_pose.shape[0])
            if betas.shape[0] != b:
                betas = betas.expand(int(b / betas.shape[0]), -1)
            vertices, joints = ref_lbs.lbs(betas, full_pose, self.v_template, self.shapedirs,
                                           self.posedirs, self.J_regressor, self.parents,
                                           self.lbs_weights, pose2rot=pose2rot)
            joints = torch.cat([joints, torch.index_select(vertices, 1, self.extra_joints_idxs)], 1)
            if transl is not None:
                joints = joints + transl.unsqueeze(1)
                vertices = vertices + transl.unsqueeze(1)
            return out_t(vertices=vertices, joints=joints, full_pose=full_pose, betas=betas,
                         global_orient=global_orient, body_pose=body_pose)

    smplx = types.ModuleType('smplx')
    smplx.SMPL = SMPL
    smplx.SMPLX = SMPL
    smplx.create = lambda *a, **k: SMPL(*a, **k)
    bm = types.ModuleType('smplx.body_models')
    bm.ModelOutput = out_t
    bm.SMPLOutput = out_t
    lb = types.ModuleType('smplx.lbs')
    lb.vertices2joints = ref_lbs.vertices2joints
    lb.lbs = ref_lbs.lbs
    smplx.body_models, smplx.lbs = bm, lb
    sys.modules['smplx'] = smplx
    sys.modules['smplx.body_models'] = bm
    sys.modules['smplx.lbs'] = lb


def setup(sm_np):
    """Make `import nemo.neural_motion_model` (the reference's) work.  Idempotent."""
    if _state.get('ready'):
        return _state['mods']
    assert available(), 'reference tree not mounted'
    if REF not in sys.path:
        sys.path.insert(0, REF)
    _install_stubs()
    _install_smplx_shim(sm_np)
    tmp = tempfile.mkdtemp(prefix='nemo_ref_')
    os.makedirs(os.path.join(tmp, 'software', 'spin_data'))
    np.save(os.path.join(tmp, 'software', 'spin_data', 'J_regressor_extra.npy'),
            sm_np['J_regressor_extra'])
    _state['cwd'] = os.getcwd()
    os.chdir(tmp)
    import warnings
    warnings.filterwarnings('ignore')
    import hmr.geometry as hg
    import utils.geometry as ug
    import hmr.smpl as hs
    import human_body_prior.body_model.lbs as rl
    import human_body_prior.models.vposer_model as vpm
    import human_body_prior.tools.rotation_tools as rt
    import hmr.smplify.prior as pr
    import nemo.rbf as rbf
    import monotonic_network as mono
    import nemo.neural_motion_model as nmm
    import nemo.utils.misc_utils as mu
    mods = types.SimpleNamespace(hg=hg, ug=ug, hs=hs, lbs=rl, vpm=vpm, rt=rt, prior=pr, rbf=rbf,
                                 mono=mono, nmm=nmm, mu=mu)
    _state.update(ready=True, mods=mods)
    return mods


def build_vposer(mods, vp_np):
    ps = types.SimpleNamespace(model_params=types.SimpleNamespace(num_neurons=512, latentD=32))
    vp = mods.vpm.VPoser(ps)
    vp.load_state_dict({k: torch.tensor(v) for k, v in vp_np.items()}, strict=False)
    for p in vp.parameters():
        p.requires_grad = False
    vp.eval()
    return vp


def build_prior(mods, gmm_np):
    pr = mods.prior.MaxMixturePrior.__new__(mods.prior.MaxMixturePrior)
    nn.Module.__init__(pr)
    pr.num_gaussians, pr.use_merged, pr.epsilon = 8, True, 1e-16
    for k in ['means', 'precisions', 'nll_weights']:
        pr.register_buffer(k, torch.tensor(gmm_np[k]))
    return pr


def build_fit_model(mods, args, params_np, sm_np, vp_np, gmm_np, data_np, version=3):
    """The reference's NemoV{2,3,4} with synthetic assets (recipe step 4)."""
    nmm = mods.nmm
    cls = {1: nmm.NemoV1, 2: nmm.NemoV2, 3: nmm.NemoV3, 4: nmm.NemoV4}[version]
    m = cls.__new__(cls)
    nn.Module.__init__(m)
    V = params_np['learned_cameras'].shape[0]
    F_ = data_np['points2d_gt_all'].shape[1]
    m.args = args
    m.device = 'cpu'
    m.FOCAL_LENGTH = 5000.0
    m.IMG_D0, m.IMG_D1 = data_np['IMG_D0'], data_np['IMG_D1']
    m.n_joints = 23
    m.num_views, m.num_frames = V, F_
    m.multi_view_seqs = types.SimpleNamespace(num_frames=F_, num_views=V)
    m.smpl = mods.hs.SMPL('unused', batch_size=1, create_transl=False)
    m.criterion_keypoints = nn.MSELoss(reduction='none')
    m.robustifier = mods.mu.GMoF()
    m.vp = build_vposer(mods, vp_np)
    m.l1_loss = nn.L1Loss(reduction='mean')
    m.pose_prior = build_prior(mods, gmm_np)
    m.points2d_gt_all = torch.tensor(data_np['points2d_gt_all'])
    m.gt_bbox_size = torch.tensor(data_np['gt_bbox_size'])
    m.hmr_theta = torch.tensor(data_np['hmr_theta'])
    m.hmr_mask = torch.tensor(data_np['hmr_mask'])
    m.learned_cameras = nn.Parameter(torch.tensor(params_np['learned_cameras']))
    if args.instance_code_size > 0:
        m.learned_instance_code = nn.Parameter(torch.tensor(params_np['learned_instance_code']))
    in_dim = (args.phase_rbf_dim if args.phase_rbf_dim > 0 else 1) + args.instance_code_size
    m.learned_motion = nmm.MotionNet(in_dim, args.h_dim, 24, 3, init_last_layer_zero=True)
    m.learned_betas = nn.Parameter(torch.tensor(params_np['learned_betas']))
    m.phase_networks = nn.ModuleList([mods.mono.MonotonicNetwork(args.monotonic_network_n_nodes, args.phase_init)
                                      for _ in range(V)])
    if args.phase_rbf_dim > 0:
        m.phase_rbf = mods.rbf.RBF(args.phase_rbf_dim, args.rbf_kernel)
    sd = {k: torch.tensor(v) for k, v in params_np.items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.split('.')[0] in ('smpl', 'vp', 'pose_prior') for k in missing), missing
    cls_opt = torch.optim.Adam if args.opt_human == 'adam' else torch.optim.AdamW
    m.opt_cameras = torch.optim.Adam([m.learned_cameras], lr=args.lr_camera, weight_decay=0)
    motion = list(m.learned_motion.parameters())
    if args.phase_rbf_dim > 0:
        motion += list(m.phase_rbf.parameters())
    m.opt_motion = cls_opt(motion, lr=args.lr_human, weight_decay=args.wd_human)
    m.opt_phase = torch.optim.Adam(m.phase_networks.parameters(), lr=args.lr_phase, weight_decay=0.0)
    m.optimizers = [m.opt_cameras, m.opt_motion, m.opt_phase]
    if args.instance_code_size > 0:
        m.opt_instance = torch.optim.Adam([m.learned_instance_code], lr=args.lr_instance, weight_decay=0)
        m.optimizers.append(m.opt_instance)
    m.detach_articulation = False
    m.start_global_traj_anywhere = False
    m.training = False
    return m
