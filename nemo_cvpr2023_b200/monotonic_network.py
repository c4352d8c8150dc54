"""Mirror of the reference's top-level `monotonic_network.py` (per-video phase warp)."""
import torch
import torch.nn as nn

from . import functional as Fn


class MonotonicNetwork(nn.Module):
    def __init__(self, n_nodes, init='rand', device='cuda'):
        super().__init__()
        self.n_nodes = n_nodes
        if init == 'linear':
            shifts_init = torch.linspace(0, 1, n_nodes)
        elif init == 'rand':
            shifts_init = torch.rand(n_nodes)
        else:
            raise ValueError()
        shifts_init.clamp_(0, 1)
        self.shifts = nn.Parameter(shifts_init.to(device))
        self.scales = nn.Parameter((torch.ones(n_nodes) * 15).to(device))

    def forward(self, x):
        return Fn.monotonic(x, self.shifts, self.scales)
