"""Mirror of `human_body_prior/body_model/body_model.py` BodyModel for model_type 'smpl'
(:182-283): axis-angle inputs, `lbs` with the skew Rodrigues, translation added to both outputs."""
import torch
import torch.nn as nn

from ... import functional as Fn
from ...hmr.smpl import load_smpl_arrays


class BodyModel(nn.Module):
    def __init__(self, bm_fname='synthetic', num_betas=10, device='cuda', **kwargs):
        super().__init__()
        self.consts = Fn.SMPLConstants(load_smpl_arrays(bm_fname), device)
        self.num_betas = num_betas
        self.model_type = 'smpl'
        self.f = torch.as_tensor(self.consts.faces, dtype=torch.int32)

    def forward(self, root_orient=None, pose_body=None, pose_hand=None, pose_jaw=None, pose_eye=None, betas=None,
                trans=None, dmpls=None, expression=None, v_template=None, joints=None, v_shaped=None,
                return_dict=False, **kwargs):
        C, dev = self.consts, self.consts.device
        B = 1
        for a in (root_orient, pose_body, pose_hand, betas, trans):
            if a is not None:
                B = a.shape[0]
                break
        z = lambda n: torch.zeros(B, n, device=dev)
        root_orient = z(3) if root_orient is None else root_orient
        pose_body = z(63) if pose_body is None else pose_body
        pose_hand = z(6) if pose_hand is None else pose_hand
        betas = z(self.num_betas) if betas is None else betas
        full_pose = torch.cat([root_orient, pose_body, pose_hand], dim=-1)
        rot = Fn.batch_rodrigues(full_pose.reshape(-1, 3), kind=1).view(B, 24, 3, 3)
        V, JP, _ = Fn.lbs_rotmats(rot, betas, C, readout=C.no_readout, want_vertices=True)
        if trans is not None:
            V, JP = V + trans.unsqueeze(1), JP + trans.unsqueeze(1)
        res = dict(v=V, f=self.f, Jtr=JP, full_pose=full_pose)
        if return_dict:
            return res

        class result_meta(object):
            pass
        out = result_meta()
        for k, v in res.items():
            setattr(out, k, v)
        return out
