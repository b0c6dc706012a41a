"""Oracle restatement of the smplx SMPL forward used by the reference (test infrastructure only).

``smplx`` is a third-party dependency of the reference (requirements.txt:22, UNPINNED, not
vendored, not installable offline), reached only through reference model/smpl.py:27
(``super().forward``).  PARITY UNPINNED: this file restates the published upstream algorithm
(smplx/lbs.py::{lbs, blend_shapes, vertices2joints, batch_rodrigues, batch_rigid_transform},
smplx/body_models.py::SMPL.forward, smplx/vertex_joint_selector.py) and is
anchored on the reference call sites (model/smpl.py:18-37, base_adaptor.py:183-190,
dynaboa_benchmark.py:216-244) and on analytic properties (tests/test_oracle.py).
"""
from types import SimpleNamespace

import torch


def smplx_rodrigues(aa):
    """smplx/lbs.py::batch_rodrigues -- R = I + sin(a) K + (1-cos(a)) K^2, a = ||aa + 1e-8||."""
    angle = (aa + 1e-8).norm(dim=1, keepdim=True)
    k = aa / angle
    zeros = torch.zeros_like(k[:, 0])
    K = torch.stack([zeros, -k[:, 2], k[:, 1], k[:, 2], zeros, -k[:, 0], -k[:, 1], k[:, 0], zeros], dim=1).view(-1, 3, 3)
    sin, cos = torch.sin(angle).unsqueeze(-1), torch.cos(angle).unsqueeze(-1)
    eye = torch.eye(3, dtype=aa.dtype).unsqueeze(0)
    return eye + sin * K + (1 - cos) * torch.bmm(K, K)


def lbs(betas, full_pose, model, pose2rot):
    """smplx/lbs.py::lbs.  ``model`` holds v_template (V,3), shapedirs (V,3,10), posedirs (207,3V),
    J_regressor (24,V), parents (24,), lbs_weights (V,24).  Returns (verts (B,V,3), J_transformed (B,24,3))."""
    B = betas.shape[0]
    dt = betas.dtype
    v_shaped = model['v_template'].unsqueeze(0) + torch.einsum('bl,mkl->bmk', betas, model['shapedirs'])
    J = torch.einsum('bik,ji->bjk', v_shaped, model['J_regressor'])
    if pose2rot:
        R = smplx_rodrigues(full_pose.reshape(-1, 3)).view(B, 24, 3, 3)
    else:
        R = full_pose.reshape(B, 24, 3, 3)
    ident = torch.eye(3, dtype=dt)
    pose_feature = (R[:, 1:] - ident).reshape(B, -1)
    v_posed = v_shaped + torch.matmul(pose_feature, model['posedirs']).view(B, -1, 3)

    # batch_rigid_transform
    parents = model['parents']
    rel = J.clone()
    rel[:, 1:] = rel[:, 1:] - J[:, parents[1:]]
    T_local = torch.zeros(B, 24, 4, 4, dtype=dt)
    T_local[:, :, :3, :3] = R
    T_local[:, :, :3, 3] = rel
    T_local[:, :, 3, 3] = 1.0
    chain = [T_local[:, 0]]
    for j in range(1, 24):
        chain.append(torch.matmul(chain[int(parents[j])], T_local[:, j]))
    G = torch.stack(chain, dim=1)
    J_transformed = G[:, :, :3, 3]
    J_h = torch.cat([J, torch.zeros(B, 24, 1, dtype=dt)], dim=2).unsqueeze(-1)
    corr = torch.matmul(G, J_h)                                     # (B,24,4,1)
    A = G - torch.nn.functional.pad(corr, [3, 0, 0, 0, 0, 0, 0, 0])

    W = model['lbs_weights'].unsqueeze(0).expand(B, -1, -1)
    T = torch.matmul(W, A.view(B, 24, 16)).view(B, -1, 4, 4)
    v_h = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=dt)], dim=2)
    verts = torch.matmul(T, v_h.unsqueeze(-1))[:, :, :3, 0]
    return verts, J_transformed


def smpl_forward(model, extra_regressor, joint_map, vertex_ids, betas, body_pose, global_orient, pose2rot=True):
    """smplx SMPL.forward + reference model/smpl.py:25-37 wrapper.

    rotmat inputs: global_orient (B,1,3,3), body_pose (B,23,3,3); axis-angle: (B,3), (B,69).
    Returns namespace(vertices (B,6890,3), joints (B,49,3), ...) like ``SMPLOutput``.
    """
    full_pose = torch.cat([global_orient, body_pose], dim=1)
    verts, J_tr = lbs(betas, full_pose, model, pose2rot)
    joints45 = torch.cat([J_tr, verts[:, vertex_ids]], dim=1)                     # vertex_joint_selector
    extra = torch.einsum('bik,ji->bjk', verts, extra_regressor)                  # vertices2joints
    joints = torch.cat([joints45, extra], dim=1)[:, joint_map]
    return SimpleNamespace(vertices=verts, joints=joints, betas=betas, global_orient=global_orient,
                           body_pose=body_pose, full_pose=full_pose)


class SMPLRef(torch.nn.Module):
    """nn.Module facade with the constructor/forward signature of reference model/smpl.py::SMPL so it
    can stand in for ``smplx.SMPL`` subclasses when the reference's own code is driven on CPU."""

    def __init__(self, model_np, extra_regressor_np, joint_map, vertex_ids, dtype=torch.float32):
        super().__init__()
        for k in ('v_template', 'shapedirs', 'posedirs', 'J_regressor', 'lbs_weights'):
            self.register_buffer(k, torch.as_tensor(model_np[k], dtype=dtype))
        self.register_buffer('parents', torch.as_tensor(model_np['parents'], dtype=torch.long))
        self.register_buffer('J_regressor_extra', torch.as_tensor(extra_regressor_np, dtype=dtype))
        self.joint_map = torch.as_tensor(joint_map, dtype=torch.long)
        self.vertex_ids = torch.as_tensor(vertex_ids, dtype=torch.long)
        self.faces = model_np['faces']

    def _model(self):
        return {k: getattr(self, k) for k in ('v_template', 'shapedirs', 'posedirs', 'J_regressor',
                                              'parents', 'lbs_weights')}

    def forward(self, betas=None, body_pose=None, global_orient=None, pose2rot=True, **_):
        return smpl_forward(self._model(), self.J_regressor_extra, self.joint_map, self.vertex_ids,
                            betas, body_pose, global_orient, pose2rot)
