"""Rotation and camera math on the CUDA library (drop-in for the hot-path functions of reference
utils/geometry.py: batch_rodrigues :9-23, rot6d_to_rotmat :47-61, perspective_projection :63-91,
rotation_matrix_to_angle_axis :184-213).  Each is a ``torch.autograd.Function`` over a hand-written
forward/adjoint kernel pair (csrc/rotmath.cuh)."""
import torch

from . import _lib, constants
from ._lib import ptr, stream


class _Rot6d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous().float()
        n = x.numel() // 6
        R = torch.empty(n, 3, 3, dtype=torch.float32, device=x.device)
        _lib.call('dboa_rot6d_fwd', ptr(x), ptr(R), n, stream())
        ctx.save_for_backward(x)
        return R

    @staticmethod
    def backward(ctx, dR):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        _lib.call('dboa_rot6d_bwd', ptr(x), ptr(dR.contiguous().float()), ptr(dx), x.numel() // 6, stream())
        return dx


def rot6d_to_rotmat(x):
    """(B,6)-like -> (N,3,3), Gram-Schmidt; accepts any shape whose numel is a multiple of 6, like the
    reference's ``x.view(-1,3,2)``."""
    _lib.require_cuda(x)
    shape = x.shape
    out = _Rot6d.apply(x.reshape(-1, 6))
    return out.view(-1, 3, 3) if len(shape) else out


class _RotmatToAA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, R):
        R = R.contiguous().float()
        n = R.numel() // 9
        aa = torch.empty(n, 3, dtype=torch.float32, device=R.device)
        _lib.call('dboa_rotmat_to_aa_fwd', ptr(R), ptr(aa), n, stream())
        ctx.save_for_backward(R)
        return aa

    @staticmethod
    def backward(ctx, daa):
        (R,) = ctx.saved_tensors
        dR = torch.empty_like(R)
        _lib.call('dboa_rotmat_to_aa_bwd', ptr(R), ptr(daa.contiguous().float()), ptr(dR), R.numel() // 9, stream())
        return dR


def rotation_matrix_to_angle_axis(rotation_matrix):
    """(N,3,3) (or (N,3,4), whose last column is ignored as in the reference) -> (N,3)."""
    _lib.require_cuda(rotation_matrix)
    if rotation_matrix.shape[-2:] == (3, 4):
        rotation_matrix = rotation_matrix[..., :3]
    return _RotmatToAA.apply(rotation_matrix.reshape(-1, 3, 3))


def batch_rodrigues(theta):
    """(N,3) axis-angle -> (N,3,3) through the quaternion route of the reference.  Forward only: the
    reference applies it to ground-truth poses (base_adaptor.py:360)."""
    _lib.require_cuda(theta)
    if theta.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError('batch_rodrigues has no backward kernel (ground-truth path only)')
    theta = theta.reshape(-1, 3).contiguous().float()
    R = torch.empty(theta.shape[0], 3, 3, dtype=torch.float32, device=theta.device)
    _lib.call('dboa_rodrigues', ptr(theta), ptr(R), theta.shape[0], 0, stream())
    return R


class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cam, j3d):
        cam, j3d = cam.contiguous().float(), j3d.contiguous().float()
        B, NJ = j3d.shape[0], j3d.shape[1]
        p2d = torch.empty(B, NJ, 2, dtype=torch.float32, device=j3d.device)
        _lib.call('dboa_project_fwd', ptr(cam), ptr(j3d), ptr(p2d), B, NJ, stream())
        ctx.save_for_backward(cam, j3d)
        return p2d

    @staticmethod
    def backward(ctx, dp):
        cam, j3d = ctx.saved_tensors
        dj, dc = torch.empty_like(j3d), torch.empty_like(cam)
        _lib.call('dboa_project_bwd', ptr(cam), ptr(j3d), ptr(dp.contiguous().float()), ptr(dj), ptr(dc), j3d.shape[0], j3d.shape[1],
                  0, 0, stream())
        return dc, dj


def project_normalized(cam, s3d):
    """``BaseAdaptor.projection(...)['normed']`` (reference base_adaptor.py:160-170): weak-perspective camera
    (s,tx,ty) -> translation (tx,ty,2f/(res*s+1e-9)), pinhole projection with f=5000, divided by res/2."""
    _lib.require_cuda(cam, s3d)
    return _Project.apply(cam, s3d)


def perspective_projection(points, rotation, translation, focal_length, camera_center):
    """General form of reference utils/geometry.py:63-91.  The hot path only ever calls it with identity
    rotation, zero camera centre and the translation of ``BaseAdaptor.projection``; that case is the fused
    kernel above.  Other argument values are rejected rather than silently computed elsewhere."""
    B = points.shape[0]
    eye = torch.eye(3, device=points.device).expand(B, 3, 3)
    if not (torch.equal(rotation, eye) and float(torch.as_tensor(camera_center).abs().max()) == 0.0
            and float(focal_length) == constants.FOCAL_LENGTH):
        raise NotImplementedError('only the BaseAdaptor.projection camera (R=I, c=0, f=5000) is implemented')
    # invert translation -> weak-perspective cam so the same kernel serves both entry points
    s = (2 * constants.FOCAL_LENGTH / translation[:, 2] - 1e-9) / constants.IMG_RES
    cam = torch.stack([s, translation[:, 0], translation[:, 1]], dim=-1)
    return project_normalized(cam, points) * (constants.IMG_RES / 2.0)
