"""Host-side mirror of tuch/smplify/prior.py (MaxMixturePrior :36-167, merged max-mixture form
:117-132) over the sm_100a pose-prior kernel (tuch_pose_terms): value and gradient in one launch.

Same constructor (`MaxMixturePrior(prior_folder=, num_gaussians=, dtype=, epsilon=, use_merged=)`),
buffers (means, covs, precisions, nll_weights, weights) and call signature `prior(pose, betas)`.
`gmm=` accepts the mixture as a dict instead of reading `gmm_{num_gaussians:02d}.pkl`.
"""
import os
import pickle

import numpy as np
import torch
import torch.nn as nn

from .. import ops


class _PriorFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, pose):
        out = ops.pose_terms(module._handle(pose.device), pose, None, pose_prior_weight=1.0,
                             want_grad=pose.requires_grad)
        ctx.save_for_backward(out['g_pose'])
        return out['prior']

    @staticmethod
    def backward(ctx, g):
        (gp,) = ctx.saved_tensors
        return None, gp * g.unsqueeze(1)


class MaxMixturePrior(nn.Module):
    def __init__(self, prior_folder='prior', num_gaussians=6, dtype=torch.float32, epsilon=1e-16,
                 use_merged=True, gmm=None, **kwargs):
        super().__init__()
        if dtype != torch.float32:
            raise ops.TuchError('MaxMixturePrior: the sm_100a kernels compute in fp32 (got dtype=%s)' % (dtype,))
        if not use_merged:
            raise ops.TuchError('MaxMixturePrior: use_merged=False (prior.py:134-161) is dead code in the '
                                'reference and is not implemented')
        self.num_gaussians = num_gaussians
        self.epsilon = epsilon
        self.use_merged = use_merged
        if gmm is None:
            path = os.path.join(prior_folder, 'gmm_{:02d}.pkl'.format(num_gaussians))
            if not os.path.exists(path):
                raise ops.TuchError('The path to the mixture prior "%s" does not exist' % path)
            with open(path, 'rb') as f:
                gmm = pickle.load(f, encoding='latin1')
        if isinstance(gmm, dict):
            means, covs, weights = gmm['means'], gmm['covars'], gmm['weights']
        elif hasattr(gmm, 'means_'):
            means, covs, weights = gmm.means_, gmm.covars_, gmm.weights_
        else:
            raise ops.TuchError('Unknown type for the prior: %s' % (type(gmm),))
        covs64 = np.asarray(covs, np.float64)
        means = np.asarray(means).astype(np.float32)
        covs = np.asarray(covs).astype(np.float32)
        # precision matrices are inverted in fp32, as the reference does (prior.py:82-83)
        precisions = np.stack([np.linalg.inv(c) for c in covs]).astype(np.float32)
        sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in covs64])
        const = (2 * np.pi) ** (69 / 2.)
        nll_weights = np.asarray(np.asarray(weights, np.float64) / (const * (sqrdets / sqrdets.min())))
        self.register_buffer('means', torch.tensor(means, dtype=dtype))
        self.register_buffer('covs', torch.tensor(covs, dtype=dtype))
        self.register_buffer('precisions', torch.tensor(precisions, dtype=dtype))
        self.register_buffer('nll_weights', torch.tensor(nll_weights, dtype=dtype).unsqueeze(dim=0))
        self.register_buffer('weights', torch.tensor(np.asarray(weights), dtype=dtype).unsqueeze(dim=0))
        self.random_var_dim = self.means.shape[1]
        self._handles = {}

    def get_mean(self):
        return torch.matmul(self.weights, self.means)

    def _handle(self, device):
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        h = self._handles.get(key)
        if h is None:
            h = ops.PriorHandle(self.means, self.precisions, self.nll_weights, device)
            self._handles[key] = h
        return h

    def merged_log_likelihood(self, pose, betas):
        return _PriorFunction.apply(self, pose)

    def forward(self, pose, betas):
        return self.merged_log_likelihood(pose, betas)
