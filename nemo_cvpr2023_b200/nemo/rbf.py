"""Mirror of `nemo/rbf.py`: RBF layer with fixed centres linspace(0,1,D) and learnable log_sigmas."""
import torch
import torch.nn as nn

from .. import functional as Fn


class RBF(nn.Module):
    def __init__(self, out_features, basis_func='linear', device='cuda'):
        super().__init__()
        assert out_features > 2
        name = basis_func if isinstance(basis_func, str) else basis_func.__name__
        if name not in Fn.RBF_BASES:
            raise ValueError('unknown basis function %r' % (name,))
        self.in_features, self.out_features, self.basis_name = 1, out_features, name
        self.register_buffer('centres', torch.linspace(0, 1, out_features, device=device).unsqueeze(1))
        self.log_sigmas = nn.Parameter(torch.zeros(out_features, device=device))

    def reset_parameters(self):
        nn.init.constant_(self.log_sigmas, 0)

    def forward(self, input):
        return Fn.rbf(input, self.centres, self.log_sigmas, self.basis_name)


def basis_func_dict():
    return {n.replace('_', ' '): n for n in Fn.RBF_BASES}
