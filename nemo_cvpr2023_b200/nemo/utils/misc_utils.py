"""Mirror of the helpers of `nemo/utils/misc_utils.py` that the fit path uses."""
import torch
import torch.nn as nn


class GMoF(nn.Module):
    """nemo/utils/misc_utils.py:91-105.  The fused fit step evaluates this inside the keypoint-loss
    kernel; the module form is kept for API parity (plain tensor expression, any device)."""

    def __init__(self, rho=100):
        super().__init__()
        self.rho = rho

    def extra_repr(self):
        return 'rho = {}'.format(self.rho)

    def forward(self, residual, sqrt):
        squared_res = residual ** 2
        if sqrt:
            squared_res = torch.sqrt(squared_res.sum(-1)).unsqueeze(-1)
        return self.rho ** 2 * torch.div(squared_res, squared_res + self.rho ** 2)


def ravel_first_2dims(x):
    return x.reshape([x.shape[0] * x.shape[1]] + list(x.shape[2:]))


def to_np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else x
