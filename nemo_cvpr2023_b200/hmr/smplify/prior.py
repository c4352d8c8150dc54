"""Mirror of `hmr/smplify/prior.py` MaxMixturePrior (merged log-likelihood, :181-196)."""
import os
import pickle
import sys

import numpy as np
import torch
import torch.nn as nn

from ... import functional as Fn


class MaxMixturePrior(nn.Module):
    def __init__(self, prior_folder='prior', num_gaussians=6, dtype=torch.float32, epsilon=1e-16, use_merged=True,
                 gmm=None, device='cuda', **kwargs):
        super().__init__()
        self.num_gaussians, self.epsilon, self.use_merged = num_gaussians, epsilon, use_merged
        if gmm is None:
            fn = os.path.join(prior_folder, 'gmm_{:02d}.pkl'.format(num_gaussians))
            if not os.path.exists(fn):
                print('The path to the mixture prior "{}"'.format(fn) + ' does not exist, exiting!')
                sys.exit(-1)
            with open(fn, 'rb') as f:
                g = pickle.load(f, encoding='latin1')
            covs = np.asarray(g['covars'], np.float64)
            sqrdets = np.sqrt(np.linalg.det(covs))
            gmm = dict(means=g['means'], precisions=np.linalg.inv(covs),
                       nll_weights=(g['weights'] / ((2 * np.pi) ** (69 / 2.) * (sqrdets / sqrdets.min())))[None])
        t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=device)
        self.register_buffer('means', t(gmm['means']))
        self.register_buffer('precisions', t(gmm['precisions']))
        self.register_buffer('nll_weights', t(gmm['nll_weights']).reshape(1, -1))
        self.random_var_dim = self.means.shape[1]

    def merged_log_likelihood(self, pose, betas):
        return Fn.gmm_prior(pose, self.means, self.precisions, self.nll_weights)

    def forward(self, pose, betas):
        return self.merged_log_likelihood(pose, betas)
