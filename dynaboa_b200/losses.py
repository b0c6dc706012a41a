"""Adaptation losses on the CUDA library, as autograd Functions whose backward was computed in the same
kernel pass as the forward (``dboa_loss_multi`` / ``dboa_loss_motion`` / ``dboa_pose_prior``).

Term order of the multi-term head (weights ``w[8]``, see include/dynaboa_b200.h):
  0 masked 2D keypoint MSE on joints 25..48   (reference base_adaptor.py:234,283,365)
  1 shape prior  mean_b sum beta^2             (:401-402)
  2 GMM pose prior                             (:405-409)
  3/4/5/6 MSE to target p2d / j3d / beta / R   (:331-337 teacher, :361-362 labelled)
  7 hip-centred masked 3D joint MSE            (:412-422)
"""
import torch

from . import _lib
from ._lib import ptr, stream

TERM_NAMES = ('s2d', 'shape_prior', 'pose_prior', 't_p2d', 't_j3d', 't_beta', 't_R', 's3d')


def _c(t):
    return None if t is None else t.detach().contiguous().float()


class _LossMulti(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p2d, j3d, R, beta, cfg):
        B, dev = p2d.shape[0], p2d.device
        p2d_c, j3d_c, R_c, beta_c = _c(p2d), _c(j3d), _c(R), _c(beta)
        w = list(cfg['w'])
        keep = [p2d_c, j3d_c, R_c, beta_c]
        dp2d, dj3d = torch.empty_like(p2d_c), torch.empty_like(j3d_c)
        dR, dbeta = torch.empty_like(R_c), torch.empty_like(beta_c)
        terms = torch.empty(9, dtype=torch.float32, device=dev)
        prior_b = None
        if w[2] != 0.0:
            prior = cfg['prior']
            prior_b = torch.empty(B, dtype=torch.float32, device=dev)
            _lib.call('dboa_pose_prior', ptr(R_c), ptr(prior.means), ptr(prior.precisions), ptr(prior.neg_log_weights), ptr(prior_b),
                      ptr(dR), float(w[2]) / B, B, stream())
        a = _lib.LossArgsStruct()
        a.B = B
        for name, t in (('p2d', p2d_c), ('j3d', j3d_c), ('R', R_c), ('beta', beta_c), ('kp', _c(cfg.get('kp'))), ('prior_b', prior_b),
                        ('t_p2d', _c(cfg.get('t_p2d'))), ('t_j3d', _c(cfg.get('t_j3d'))), ('t_beta', _c(cfg.get('t_beta'))),
                        ('t_R', _c(cfg.get('t_R'))), ('gt_s3d', _c(cfg.get('gt_s3d'))), ('terms', terms), ('dp2d', dp2d),
                        ('dj3d', dj3d), ('dR', dR), ('dbeta', dbeta)):
            keep.append(t)
            setattr(a, name, None if t is None else t.data_ptr())
        for i in range(8):
            a.w[i] = float(w[i])
        a.dR_accumulate = 1 if prior_b is not None else 0
        import ctypes as C
        _lib.call('dboa_loss_multi', C.byref(a), stream())
        ctx.save_for_backward(dp2d, dj3d, dR, dbeta)
        ctx.shapes = (p2d.shape, j3d.shape, R.shape, beta.shape)
        total, parts = terms[8], terms[:8]
        ctx.mark_non_differentiable(parts)
        return total, parts

    @staticmethod
    def backward(ctx, g, _):
        dp2d, dj3d, dR, dbeta = ctx.saved_tensors
        s = ctx.shapes
        return (dp2d * g).view(s[0]), (dj3d * g).view(s[1]), (dR * g).view(s[2]), (dbeta * g).view(s[3]), None


def loss_multi(p2d, j3d, R, beta, weights, prior=None, kp=None, t_p2d=None, t_j3d=None, t_beta=None, t_R=None, gt_s3d=None):
    """Weighted sum of the selected terms; returns (total, terms[8]) with gradients to (p2d, j3d, R, beta)."""
    _lib.require_cuda(p2d, j3d, R, beta)
    if weights[2] != 0.0 and prior is None:
        raise ValueError('pose prior weight set but no MaxMixturePrior given')
    cfg = dict(w=weights, prior=prior, kp=kp, t_p2d=t_p2d, t_j3d=t_j3d, t_beta=t_beta, t_R=t_R, gt_s3d=gt_s3d)
    return _LossMulti.apply(p2d, j3d, R.reshape(-1, 24, 3, 3), beta, cfg)


class _LossMotion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p_cur, p_hist, kp_cur, kp_hist):
        B = p_cur.shape[0]
        pc, ph = _c(p_cur), _c(p_hist)
        d_cur, d_hist = torch.empty_like(pc), torch.empty_like(ph)
        term = torch.empty(1, dtype=torch.float32, device=pc.device)
        _lib.call('dboa_loss_motion', ptr(pc), ptr(ph), ptr(_c(kp_cur)), ptr(_c(kp_hist)), 1.0, ptr(term), ptr(d_cur), ptr(d_hist), B, 0,
                  stream())
        ctx.save_for_backward(d_cur, d_hist)
        return term[0]

    @staticmethod
    def backward(ctx, g):
        d_cur, d_hist = ctx.saved_tensors
        return d_cur * g, d_hist * g, None, None


def loss_motion(p_cur, p_hist, kp_cur, kp_hist):
    """mean over (B,24,2) of [both visible] * ((p_cur - p_hist) - (kp_cur - kp_hist))^2 on joints 25..48
    (reference base_adaptor.py:387-396).  All tensors are full 49-joint arrays."""
    _lib.require_cuda(p_cur, p_hist, kp_cur, kp_hist)
    return _LossMotion.apply(p_cur, p_hist, kp_cur, kp_hist)
