"""Oracle restatement of reference utils/geometry.py (test infrastructure only).

Every function names the reference lines it follows.  Plain torch on CPU; differentiable
through torch autograd, which is how gradient parity of the CUDA backward kernels is checked.
"""
import torch


def _unit(v, eps=1e-12):
    """F.normalize(v, dim=1): v / max(||v||, eps)."""
    return v / v.norm(dim=1, keepdim=True).clamp_min(eps)


def rot6d_to_rotmat(x):
    """reference utils/geometry.py:47-61 -- Gram-Schmidt on the two interleaved 3-vectors."""
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = _unit(a1)
    b2 = _unit(a2 - (b1 * a2).sum(1, keepdim=True) * b1)
    b3 = torch.linalg.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def quat_to_rotmat(quat):
    """reference utils/geometry.py:25-45 -- normalise then expand (w,x,y,z) to 3x3."""
    q = quat / quat.norm(dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    rows = [w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
            2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
            2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2]
    return torch.stack(rows, dim=1).view(-1, 3, 3)


def batch_rodrigues(theta):
    """reference utils/geometry.py:9-23 -- axis-angle -> quaternion (angle = ||theta+1e-8||) -> R."""
    angle = (theta + 1e-8).norm(dim=1, keepdim=True)
    axis = theta / angle
    half = angle * 0.5
    quat = torch.cat([torch.cos(half), torch.sin(half) * axis], dim=1)
    return quat_to_rotmat(quat)


def rotation_matrix_to_quaternion(R, eps=1e-6):
    """reference utils/geometry.py:248-306 -- four-branch masked-sum quaternion extraction.

    ``R`` is (N,3,3); the reference first pads it to (N,3,4) and then transposes
    (geometry.py:203-208, :264), so ``t[:, i, j]`` below is ``R[:, j, i]``.
    """
    t = R.transpose(1, 2)
    d2 = t[:, 2, 2] < eps
    d0_gt_d1 = t[:, 0, 0] > t[:, 1, 1]
    d0_lt_nd1 = t[:, 0, 0] < -t[:, 1, 1]

    t0 = 1 + t[:, 0, 0] - t[:, 1, 1] - t[:, 2, 2]
    q0 = torch.stack([t[:, 1, 2] - t[:, 2, 1], t0, t[:, 0, 1] + t[:, 1, 0], t[:, 2, 0] + t[:, 0, 2]], -1)
    t1 = 1 - t[:, 0, 0] + t[:, 1, 1] - t[:, 2, 2]
    q1 = torch.stack([t[:, 2, 0] - t[:, 0, 2], t[:, 0, 1] + t[:, 1, 0], t1, t[:, 1, 2] + t[:, 2, 1]], -1)
    t2 = 1 - t[:, 0, 0] - t[:, 1, 1] + t[:, 2, 2]
    q2 = torch.stack([t[:, 0, 1] - t[:, 1, 0], t[:, 2, 0] + t[:, 0, 2], t[:, 1, 2] + t[:, 2, 1], t2], -1)
    t3 = 1 + t[:, 0, 0] + t[:, 1, 1] + t[:, 2, 2]
    q3 = torch.stack([t3, t[:, 1, 2] - t[:, 2, 1], t[:, 2, 0] - t[:, 0, 2], t[:, 0, 1] - t[:, 1, 0]], -1)

    c0 = (d2 & d0_gt_d1).view(-1, 1).type_as(q0)
    c1 = (d2 & ~d0_gt_d1).view(-1, 1).type_as(q0)
    c2 = (~d2 & d0_lt_nd1).view(-1, 1).type_as(q0)
    c3 = (~d2 & ~d0_lt_nd1).view(-1, 1).type_as(q0)
    q = q0 * c0 + q1 * c1 + q2 * c2 + q3 * c3
    denom = torch.sqrt(t0.view(-1, 1) * c0 + t1.view(-1, 1) * c1 + t2.view(-1, 1) * c2 + t3.view(-1, 1) * c3)
    return q / denom * 0.5


def quaternion_to_angle_axis(q):
    """reference utils/geometry.py:216-245."""
    q1, q2, q3 = q[..., 1], q[..., 2], q[..., 3]
    sin_sq = q1 * q1 + q2 * q2 + q3 * q3
    sin_t = torch.sqrt(sin_sq)
    cos_t = q[..., 0]
    two_theta = 2.0 * torch.where(cos_t < 0.0, torch.atan2(-sin_t, -cos_t), torch.atan2(sin_t, cos_t))
    k = torch.where(sin_sq > 0.0, two_theta / sin_t, 2.0 * torch.ones_like(sin_t))
    return torch.stack([q1 * k, q2 * k, q3 * k], dim=-1)


def rotation_matrix_to_angle_axis(R):
    """reference utils/geometry.py:184-213 (NaN entries are zeroed, :212)."""
    aa = quaternion_to_angle_axis(rotation_matrix_to_quaternion(R.reshape(-1, 3, 3)))
    return torch.where(torch.isnan(aa), torch.zeros_like(aa), aa)


def perspective_projection(points, rotation, translation, focal_length, camera_center):
    """reference utils/geometry.py:63-91."""
    B = points.shape[0]
    K = torch.zeros(B, 3, 3, dtype=points.dtype)
    K[:, 0, 0] = focal_length
    K[:, 1, 1] = focal_length
    K[:, 2, 2] = 1.0
    K[:, :-1, -1] = camera_center
    p = torch.einsum('bij,bkj->bki', rotation, points) + translation.unsqueeze(1)
    p = p / p[:, :, -1].unsqueeze(-1)
    p = torch.einsum('bij,bkj->bki', K, p)
    return p[:, :, :-1]


def weak_perspective_project(cam, s3d, focal=5000.0, res=224, eps=1e-9):
    """reference base_adaptor.py:160-170 (``BaseAdaptor.projection``): returns (ori, normed)."""
    B = s3d.shape[0]
    cam_t = torch.stack([cam[:, 1], cam[:, 2], 2 * focal / (res * cam[:, 0] + eps)], dim=-1)
    eye = torch.eye(3, dtype=s3d.dtype).unsqueeze(0).expand(B, -1, -1)
    s2d = perspective_projection(s3d, eye, cam_t, focal, torch.zeros(B, 2, dtype=s3d.dtype))
    return s2d, s2d / (res / 2.0)
