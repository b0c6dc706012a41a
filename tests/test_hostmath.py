"""CPU check of the hand-derived adjoints in csrc/rotmath.cuh (compiled for the host) against torch autograd
of the oracle restatements.  The same header is what the device kernels include."""
import ctypes

import numpy as np
import pytest
import torch

from dynaboa_b200 import build, constants
from oracle import geometry_ref as G, smplx_ref
from conftest import rel_err


@pytest.fixture(scope='module')
def lib():
    return ctypes.CDLL(build.build_hostmath())


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_rot6d(lib):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(100, 6, generator=g)
    x[0] = torch.tensor([1., 0, 0, 1, 0, 0])
    xr = x.clone().requires_grad_(True)
    R = G.rot6d_to_rotmat(xr)
    w = torch.randn(100, 3, 3, generator=g)
    (R * w).sum().backward()
    Rn, dx = np.zeros((100, 9), np.float32), np.zeros((100, 6), np.float32)
    lib.hm_rot6d_fwd(P(x.numpy()), P(Rn), 100)
    lib.hm_rot6d_bwd(P(x.numpy()), P(w.numpy().copy()), P(dx), 100)
    assert rel_err(Rn.reshape(100, 3, 3), R.detach()) < 1e-6
    assert rel_err(dx, xr.grad) < 1e-5


def test_rodrigues(lib):
    g = torch.Generator().manual_seed(2)
    aa = torch.randn(200, 3, generator=g) * torch.linspace(0.01, 3.1, 200).unsqueeze(1)
    aa[0] = 0
    R = np.zeros((200, 9), np.float32)
    lib.hm_quat_rodrigues(P(aa.numpy()), P(R), 200)
    assert rel_err(R.reshape(200, 3, 3), G.batch_rodrigues(aa)) < 2e-6
    lib.hm_smplx_rodrigues(P(aa.numpy()), P(R), 200)
    assert rel_err(R.reshape(200, 3, 3), smplx_ref.smplx_rodrigues(aa)) < 2e-6


def test_rotmat_to_aa_all_branches(lib):
    g = torch.Generator().manual_seed(3)
    big = torch.randn(400, 3, generator=g)
    big = big / big.norm(dim=1, keepdim=True) * torch.linspace(0.001, 3.14, 400).unsqueeze(1)
    Rm = G.batch_rodrigues(big)
    Rr = Rm.clone().requires_grad_(True)
    a = G.rotation_matrix_to_angle_axis(Rr)
    w = torch.randn(400, 3, generator=g)
    (a * w).sum().backward()
    an, dR = np.zeros((400, 3), np.float32), np.zeros((400, 9), np.float32)
    lib.hm_r2aa_fwd(P(Rm.numpy()), P(an), 400)
    lib.hm_r2aa_bwd(P(Rm.numpy()), P(w.numpy().copy()), P(dR), 400)
    assert rel_err(an, a.detach()) < 1e-6
    assert rel_err(dR.reshape(400, 3, 3), Rr.grad) < 1e-5
    t = Rm
    d2 = t[:, 2, 2] < 1e-6
    branches = {int(v) for v in ((d2 & ~(t[:, 0, 0] > t[:, 1, 1])).int() + (~d2 & (t[:, 0, 0] < -t[:, 1, 1])).int() * 2
                                 + (~d2 & ~(t[:, 0, 0] < -t[:, 1, 1])).int() * 3).tolist()}
    assert branches == {0, 1, 2, 3}


def test_projection(lib):
    g = torch.Generator().manual_seed(4)
    cam = torch.tensor([[0.9, 0.01, -0.02], [1.1, 0.1, 0.05]]).requires_grad_(True)
    X = (torch.randn(2, 49, 3, generator=g) * 0.4).requires_grad_(True)
    p = G.weak_perspective_project(cam, X)[1]
    w = torch.randn(2, 49, 2, generator=g)
    (p * w).sum().backward()
    pn, dX, dc = np.zeros((2, 49, 2), np.float32), np.zeros((2, 49, 3), np.float32), np.zeros((2, 3), np.float32)
    lib.hm_project_fwd(P(cam.detach().numpy()), P(X.detach().numpy()), P(pn), 2, 49)
    lib.hm_project_bwd(P(cam.detach().numpy()), P(X.detach().numpy()), P(w.numpy().copy()), P(dX), P(dc), 2, 49)
    assert rel_err(pn, p.detach()) < 1e-6
    assert rel_err(dX, X.grad) < 1e-5 and rel_err(dc, cam.grad) < 1e-5


def test_kinematic_chain(lib):
    g = torch.Generator().manual_seed(5)
    B = 3
    R = G.batch_rodrigues(torch.randn(B * 24, 3, generator=g) * 0.5).view(B, 24, 3, 3).clone().requires_grad_(True)
    J = (torch.randn(B, 24, 3, generator=g) * 0.3).requires_grad_(True)
    parents = torch.tensor(constants.SMPL_PARENTS)
    rel = torch.cat([J[:, :1], J[:, 1:] - J[:, parents[1:]]], 1)
    T = torch.zeros(B, 24, 4, 4)
    T[:, :, :3, :3], T[:, :, :3, 3], T[:, :, 3, 3] = R, rel, 1
    ch = [T[:, 0]]
    for j in range(1, 24):
        ch.append(ch[int(parents[j])] @ T[:, j])
    Gm = torch.stack(ch, 1)
    Jh = torch.cat([J, torch.zeros(B, 24, 1)], 2).unsqueeze(-1)
    A = (Gm - torch.nn.functional.pad(Gm @ Jh, [3, 0, 0, 0, 0, 0, 0, 0]))[:, :, :3, :]
    wA, wJ = torch.randn(B, 24, 3, 4, generator=g), torch.randn(B, 24, 3, generator=g)
    ((A * wA).sum() + (Gm[:, :, :3, 3] * wJ).sum()).backward()
    Rn, Jn = R.detach().numpy().reshape(B, 216).copy(), J.detach().numpy().reshape(B, 72).copy()
    pn = parents.numpy().astype(np.int32)
    Gr, Gt, An = np.zeros((B, 216), np.float32), np.zeros((B, 72), np.float32), np.zeros((B, 288), np.float32)
    lib.hm_chain_fwd(P(Rn), P(Jn), P(pn), P(Gr), P(Gt), P(An), B)
    assert rel_err(An.reshape(B, 24, 3, 4), A.detach()) < 1e-6
    dR, dJ = np.zeros((B, 216), np.float32), np.zeros((B, 72), np.float32)
    lib.hm_chain_bwd(P(Rn), P(Jn), P(pn), P(Gr), P(wA.numpy().reshape(B, 288).copy()), P(wJ.numpy().reshape(B, 72).copy()), P(dR), P(dJ), B)
    assert rel_err(dR.reshape(B, 24, 3, 3), R.grad) < 1e-5 and rel_err(dJ.reshape(B, 24, 3), J.grad) < 1e-5
