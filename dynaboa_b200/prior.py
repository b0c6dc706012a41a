"""Max-mixture GMM pose prior on the CUDA library (drop-in for reference utils/smplify/prior.py
``MaxMixturePrior`` :100-231, merged path :181-196).

Constants are prepared on the host exactly as the reference does (float64 determinants, fp32 cast of the
inverse covariances, ``nll_weights`` underflowing to 0 / denormals in fp32 -- SURVEY.md Appendix D) and the
kernel receives ``-log(nll_weights)`` (one entry is +inf and can never be the minimum).
"""
import os

import numpy as np
import torch
from torch import nn

from . import _lib, config
from ._lib import ptr, stream


def load_gmm(prior_folder=None, num_gaussians=8):
    """Accepts the reference's ``gmm_XX.pkl`` (in ``prior_folder``) or the packaged npz copy of it."""
    if prior_folder is not None:
        pkl = os.path.join(prior_folder, 'gmm_{:02d}.pkl'.format(num_gaussians))
        if os.path.exists(pkl):
            import pickle
            with open(pkl, 'rb') as f:
                g = pickle.load(f, encoding='latin1')
            return {k: np.asarray(g[k]) for k in ('means', 'covars', 'weights')}
        npz = os.path.join(prior_folder, 'gmm_{:02d}.npz'.format(num_gaussians))
        if os.path.exists(npz):
            return dict(np.load(npz))
    return dict(np.load(config.GMM_PRIOR))


class _GmmFromPose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, owner):
        pose = pose.contiguous().float()
        B = pose.shape[0]
        out = torch.empty(B, dtype=torch.float32, device=pose.device)
        grad = torch.empty_like(pose)
        _lib.call('dboa_gmm_prior', ptr(pose), ptr(owner.means), ptr(owner.precisions), ptr(owner.neg_log_weights), ptr(out),
                  ptr(grad), 1.0, B, stream())
        ctx.save_for_backward(grad)
        return out

    @staticmethod
    def backward(ctx, go):
        (grad,) = ctx.saved_tensors
        return grad * go.unsqueeze(1), None


class _GmmFromRotmat(torch.autograd.Function):
    """rotation_matrix_to_angle_axis + merged GMM in one kernel (reference base_adaptor.py:405-409)."""

    @staticmethod
    def forward(ctx, rotmat, owner):
        rotmat = rotmat.contiguous().float()
        B = rotmat.shape[0]
        out = torch.empty(B, dtype=torch.float32, device=rotmat.device)
        grad = torch.empty_like(rotmat)
        _lib.call('dboa_pose_prior', ptr(rotmat), ptr(owner.means), ptr(owner.precisions), ptr(owner.neg_log_weights), ptr(out),
                  ptr(grad), 1.0, B, stream())
        ctx.save_for_backward(grad)
        return out

    @staticmethod
    def backward(ctx, go):
        (grad,) = ctx.saved_tensors
        return grad * go.view(-1, 1, 1, 1), None


class MaxMixturePrior(nn.Module):
    def __init__(self, prior_folder='prior', num_gaussians=8, dtype=torch.float32, epsilon=1e-16, use_merged=True, **kwargs):
        super().__init__()
        if dtype != torch.float32 or not use_merged or num_gaussians != 8:
            raise NotImplementedError('the CUDA prior implements the reference configuration: 8 gaussians, fp32, merged')
        gmm = load_gmm(prior_folder, num_gaussians)
        means = gmm['means'].astype(np.float32)
        covs = gmm['covars'].astype(np.float32)
        precisions = np.stack([np.linalg.inv(c) for c in covs]).astype(np.float32)
        sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in gmm['covars']])
        const = (2 * np.pi) ** (69 / 2.0)
        nll_weights = torch.tensor(np.asarray(gmm['weights'] / (const * (sqrdets / sqrdets.min()))), dtype=torch.float32).unsqueeze(0)
        self.register_buffer('means', torch.tensor(means))
        self.register_buffer('covs', torch.tensor(covs))
        self.register_buffer('precisions', torch.tensor(precisions).contiguous())
        self.register_buffer('nll_weights', nll_weights)
        self.register_buffer('neg_log_weights', (-torch.log(nll_weights)).reshape(-1).contiguous())
        self.register_buffer('weights', torch.tensor(gmm['weights'], dtype=torch.float32).unsqueeze(0))
        self.num_gaussians, self.epsilon, self.use_merged = num_gaussians, epsilon, use_merged
        self.random_var_dim = 69

    def get_mean(self):
        return torch.matmul(self.weights, self.means)

    def merged_log_likelihood(self, pose, betas=None):
        _lib.require_cuda(pose, self.means)
        return _GmmFromPose.apply(pose.reshape(-1, 69), self)

    def from_rotmat(self, rotmat):
        """(B,24,3,3) rotations -> (B,) prior of the 23 body joints, fused with the axis-angle conversion."""
        _lib.require_cuda(rotmat, self.means)
        return _GmmFromRotmat.apply(rotmat.reshape(-1, 24, 3, 3), self)

    def forward(self, pose, betas=None):
        return self.merged_log_likelihood(pose, betas)
