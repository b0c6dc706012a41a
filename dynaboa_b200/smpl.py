"""SMPL body model on the CUDA library (drop-in for reference model/smpl.py).

``SMPL(model_path, gender='neutral', create_transl=False, batch_size=1)`` and
``forward(betas=, body_pose=, global_orient=, pose2rot=True)`` keep the signature of the reference wrapper
(model/smpl.py:15-37) over ``smplx.SMPL``; the output object exposes ``vertices (B,6890,3)``,
``joints (B,49,3)``, ``global_orient``, ``body_pose``, ``betas``, ``full_pose``.

Model data is read from ``<model_path>/SMPL_<GENDER>.npz`` (arrays named as smplx's buffers: v_template,
shapedirs, posedirs, J_regressor, parents, lbs_weights, faces).  The licensed SMPL pickles are not
available offline; ``dynaboa_b200.synthetic.write_asset_dir`` writes stand-ins with these names.
"""
import ctypes as C
import os
from types import SimpleNamespace

import numpy as np
import torch
from torch import nn

from . import _lib, config, constants
from ._lib import ptr, stream


def vertices2joints(J_regressor, vertices):
    """smplx.lbs.vertices2joints (used by the reference driver for the H36M regressor)."""
    return torch.einsum('bik,ji->bjk', vertices, J_regressor)


class SMPLOutput(SimpleNamespace):
    pass


class _SMPLFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, betas, rotmat, owner):
        B = betas.shape[0]
        dev = betas.device
        betas_c, rot_c = betas.contiguous().float(), rotmat.contiguous().float()
        verts = torch.empty(B, 6890, 3, dtype=torch.float32, device=dev)
        joints = torch.empty(B, 49, 3, dtype=torch.float32, device=dev)
        tape = torch.empty(_lib.load().dboa_smpl_tape_floats(B), dtype=torch.float32, device=dev)
        _lib.call('dboa_smpl_forward', owner._struct_ref(), ptr(betas_c), ptr(rot_c), B, ptr(verts), ptr(joints), ptr(tape), stream())
        ctx.owner, ctx.B = owner, B
        ctx.save_for_backward(rot_c, tape)
        ctx.mark_non_differentiable(verts)
        return verts, joints

    @staticmethod
    def backward(ctx, d_verts, d_joints):
        rot_c, tape = ctx.saved_tensors
        B, dev = ctx.B, rot_c.device
        if d_joints is None:
            return None, None, None
        d_rot = torch.empty(B, 24, 3, 3, dtype=torch.float32, device=dev)
        d_betas = torch.empty(B, 10, dtype=torch.float32, device=dev)
        scratch = torch.empty(_lib.load().dboa_smpl_scratch_floats(B), dtype=torch.float32, device=dev)
        _lib.call('dboa_smpl_backward', ctx.owner._struct_ref(), ptr(rot_c), B, ptr(tape), ptr(d_joints.contiguous().float()),
                  ptr(scratch), ptr(d_rot), ptr(d_betas), 0, stream())
        return d_betas, d_rot, None


class SMPL(nn.Module):
    """Extension of SMPL to the 49 SPIN joints (reference model/smpl.py:15-37)."""

    def __init__(self, model_path=None, gender='neutral', create_transl=False, batch_size=1, data=None,
                 extra_regressor=None, **kwargs):
        super().__init__()
        if create_transl:
            raise NotImplementedError('create_transl=True is never used by the reference (base_adaptor.py:144-146)')
        if data is None:
            data = dict(np.load(os.path.join(model_path or config.SMPL_MODEL_DIR, f'SMPL_{gender.upper()}.npz')))
        if extra_regressor is None:
            extra_regressor = np.load(config.JOINT_REGRESSOR_TRAIN_EXTRA)
        f32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).contiguous()
        v_template, shapedirs = f32(data['v_template']), f32(data['shapedirs'])
        posedirs, J_reg = f32(data['posedirs']), f32(data['J_regressor'])
        nv = v_template.shape[0]
        if nv != constants.NUM_VERTS or tuple(posedirs.shape) != (207, nv * 3) or tuple(shapedirs.shape) != (nv, 3, 10):
            raise ValueError('SMPL model arrays have unexpected shapes')
        # smplx buffers kept under their upstream names
        self.register_buffer('v_template', v_template)
        self.register_buffer('shapedirs', shapedirs)
        self.register_buffer('posedirs', posedirs)
        self.register_buffer('J_regressor', J_reg)
        self.register_buffer('lbs_weights', f32(data['lbs_weights']))
        self.register_buffer('parents', torch.as_tensor(np.asarray(data['parents']), dtype=torch.int32))
        self.register_buffer('J_regressor_extra', f32(extra_regressor))
        # kernel-side layouts (include/dynaboa_b200.h: dboa_smpl_model)
        self.register_buffer('blend_dirs', torch.cat([shapedirs.permute(2, 0, 1).reshape(10, nv * 3), posedirs], 0).contiguous())
        # J_regressor folded through the template / shape blend (done once, in float64)
        Jd = J_reg.double()
        self.register_buffer('J_template', (Jd @ v_template.double()).float().contiguous())
        self.register_buffer('J_shapedirs', torch.einsum('jv,vkl->jkl', Jd, shapedirs.double()).float().contiguous())
        self.register_buffer('joint_map_i32', torch.tensor(constants.JOINT_MAP_49, dtype=torch.int32))
        self.register_buffer('vertex_ids_i32', torch.tensor(constants.SMPL_EXTRA_VERTEX_IDS, dtype=torch.int32))
        self.joint_map = torch.tensor(constants.JOINT_MAP_49, dtype=torch.long)
        self.faces = np.asarray(data['faces']) if 'faces' in data else None
        self.gender = gender
        self._struct = None

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self._struct = None
        return self

    def _struct_ref(self):
        if self._struct is None:
            s = _lib.SmplModelStruct()
            for name, buf in (('v_template', self.v_template), ('blend_dirs', self.blend_dirs), ('J_template', self.J_template),
                              ('J_shapedirs', self.J_shapedirs), ('parents', self.parents), ('lbs_weights', self.lbs_weights),
                              ('J_extra', self.J_regressor_extra), ('joint_map', self.joint_map_i32),
                              ('vertex_ids', self.vertex_ids_i32)):
                _lib.require_cuda(buf)
                setattr(s, name, buf.data_ptr())
            self._struct = s
        return C.byref(self._struct)

    def forward(self, betas=None, body_pose=None, global_orient=None, pose2rot=True, **kwargs):
        B = betas.shape[0]
        _lib.require_cuda(betas, body_pose, global_orient)
        if pose2rot:
            full_pose = torch.cat([global_orient.reshape(B, 3), body_pose.reshape(B, 69)], dim=1)
            aa = full_pose.reshape(-1, 3).contiguous().float()
            rotmat = torch.empty(B, 24, 3, 3, dtype=torch.float32, device=betas.device)
            _lib.call('dboa_rodrigues', ptr(aa), ptr(rotmat), B * 24, 1, stream())     # smplx lbs.batch_rodrigues
        else:
            full_pose = torch.cat([global_orient.reshape(B, 1, 3, 3), body_pose.reshape(B, 23, 3, 3)], dim=1)
            rotmat = full_pose
        verts, joints = _SMPLFunction.apply(betas, rotmat, self)
        return SMPLOutput(vertices=verts, joints=joints, global_orient=global_orient, body_pose=body_pose, betas=betas,
                          full_pose=full_pose)

    def get_smpl_faces(self):
        return self.faces


def get_smpl_faces():
    """reference model/smpl.py:45-47."""
    return SMPL(config.SMPL_MODEL_DIR, batch_size=1, create_transl=False).faces
