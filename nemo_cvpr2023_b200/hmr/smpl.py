"""Mirror of the reference's `hmr/smpl.py` SMPL class (`hmr/smpl.py:17-43`), including the part of
`smplx.SMPL.forward` it inherits (smplx 0.1.28): same constructor arguments, same forward signature,
same output fields."""
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn

from .. import assets
from .. import functional as Fn

ModelOutput = namedtuple('ModelOutput', ['vertices', 'joints', 'full_pose', 'betas', 'global_orient', 'body_pose'])
ModelOutput.__new__.__defaults__ = (None,) * 6


def load_smpl_arrays(model_path):
    """dict of arrays (see assets.make_smpl_assets) from a dict, an .npz, or -- for the offline
    benchmark, where the licensed SMPL pickle is unavailable -- the seeded synthetic model."""
    if isinstance(model_path, dict):
        return model_path
    if isinstance(model_path, str) and model_path.endswith('.npz'):
        d = dict(np.load(model_path))
        base = assets.make_smpl_assets(0)
        for k in ('parents', 'vertex_joint_ids', 'joint_map', 'faces'):
            d.setdefault(k, base[k])
        return d
    if model_path in (None, 'synthetic'):
        return assets.make_smpl_assets(0)
    raise FileNotFoundError(
        'SMPL model %r: pass a dict / .npz with v_template, shapedirs, posedirs, J_regressor, lbs_weights, '
        'J_regressor_extra, or "synthetic" for the seeded stand-in (the licensed pickle cannot be shipped)' % (model_path,))


class SMPL(nn.Module):
    """Extension of SMPL with the 9 extra regressed joints and the 49-joint map (hmr/smpl.py:17-27)."""

    def __init__(self, model_path='synthetic', batch_size=1, create_transl=False, device='cuda', **kwargs):
        super().__init__()
        sm = load_smpl_arrays(model_path)
        self.consts = Fn.SMPLConstants(sm, device)
        self.batch_size = batch_size
        self.faces = self.consts.faces
        for k, v in self.consts.raw.items():
            self.register_buffer(k, v, persistent=False)
        self.register_buffer('parents', self.consts.parents.long(), persistent=False)
        self.joint_map = self.consts.joint_map.long()

    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None, return_verts=True,
                return_full_pose=False, pose2rot=True, **kwargs):
        C = self.consts
        dev = C.device
        bs = self.batch_size
        # smplx: missing arguments fall back to the (zero) model parameters
        if global_orient is None:
            global_orient = torch.zeros(bs, 3, device=dev) if pose2rot else torch.eye(3, device=dev).expand(bs, 1, 3, 3)
        if body_pose is None:
            body_pose = torch.zeros(bs, 69, device=dev) if pose2rot else torch.eye(3, device=dev).expand(bs, 23, 3, 3)
        if betas is None:
            betas = torch.zeros(bs, C.num_betas, device=dev)
        full_pose = torch.cat([global_orient, body_pose], dim=1)
        N = max(betas.shape[0], global_orient.shape[0], body_pose.shape[0])
        if betas.shape[0] != N:
            # smplx expands a batch-1 betas; the kernels broadcast it (shared-shape fast path) instead
            if betas.shape[0] != 1:
                raise ValueError('betas batch %d cannot broadcast to %d' % (betas.shape[0], N))
        if pose2rot:
            rot = Fn.batch_rodrigues(full_pose.reshape(-1, 3), kind=1).view(N, 24, 3, 3)
        else:
            rot = full_pose.reshape(N, 24, 3, 3)
        V, JP, JX = Fn.lbs_rotmats(rot, betas, C, want_vertices=True)
        ro = C.readout
        joints = Fn.joints_gather(JP, JX, C.joint_map, ro['np'], transl, ro['rowsum'])
        if transl is not None:
            V = V + transl.unsqueeze(1)
        out_betas = betas if betas.shape[0] == N else betas.expand(N, -1)
        return ModelOutput(vertices=V if return_verts else None, joints=joints, full_pose=full_pose,
                           betas=out_betas, global_orient=global_orient, body_pose=body_pose)
