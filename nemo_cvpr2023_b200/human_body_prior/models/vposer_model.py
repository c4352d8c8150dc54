"""Mirror of `human_body_prior/models/vposer_model.py` (VPoser V02_05) in eval mode, as the fit path
uses it (`model_loader.load_model` disables grads and calls .eval(), model_loader.py:65-83)."""
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from ... import functional as Fn
from ...fit_engine import fold_vposer
from ..tools.rotation_tools import matrot2aa


class ContinousRotReprDecoder(nn.Module):
    def forward(self, module_input):
        return Fn.rot6d_to_rotmat(module_input.reshape(-1, 6))


class VPoser(nn.Module):
    """encode() returns a Normal(mu, softplus(.)) like the reference (vposer_model.py:48-56, 90-96)."""

    def __init__(self, model_ps=None, state_dict=None, device='cuda'):
        super().__init__()
        ps = model_ps.model_params if model_ps is not None else SimpleNamespace(num_neurons=512, latentD=32)
        self.latentD, self.num_neurons, self.num_joints = ps.latentD, ps.num_neurons, 21
        self._dev = torch.device(device)
        self._folded = None
        if state_dict is not None:
            self.load_reference_state_dict(state_dict)
        self.eval()

    def load_reference_state_dict(self, sd):
        self._sd = {k: np.asarray(v.detach().cpu() if torch.is_tensor(v) else v) for k, v in sd.items()}
        self._folded = fold_vposer(self._sd, self._dev)

    def encode(self, pose_body):
        f = self._folded
        x = pose_body.reshape(pose_body.shape[0], -1)
        h = Fn.linear(x, f['W1'], f['b1'], 'leaky_relu')
        head = Fn.linear(h, f['Wh'], f['bh'])
        mu, pre = head[:, :self.latentD], head[:, self.latentD:]
        return torch.distributions.normal.Normal(mu, torch.nn.functional.softplus(pre))

    def decode(self, Zin):
        f = self._folded
        bs = Zin.shape[0]
        h = Fn.linear(Zin, f['Wd0'], f['bd0'], 'leaky_relu')
        h = Fn.linear(h, f['Wd3'], f['bd3'], 'leaky_relu')
        prec = Fn.rot6d_to_rotmat(Fn.linear(h, f['Wd5'], f['bd5']).reshape(-1, 6))
        return {'pose_body': matrot2aa(prec.view(-1, 3, 3)).view(bs, -1, 3), 'pose_body_matrot': prec.view(bs, -1, 9)}

    def forward(self, pose_body):
        q_z = self.encode(pose_body)
        out = self.decode(q_z.rsample())
        out.update({'poZ_body_mean': q_z.mean, 'poZ_body_std': q_z.scale, 'q_z': q_z})
        return out
