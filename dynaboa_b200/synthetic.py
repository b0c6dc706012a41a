"""Seeded synthetic assets and input stream for the DynaBOA hot path.

None of the reference's real assets exist offline (SMPL model files, basemodel.pt,
H36M exemplars, 3DPW frames: SURVEY.md §8c/§8d), so tests, ``bench.py`` and
``__graft_entry__.smoke()`` run on a deterministic stand-in with the same shapes,
dtypes and dict keys:

* ``make_smpl_model``      -- SMPL-shaped body model buffers (smplx ``SMPL`` data layout,
                              SURVEY.md Appendix A)
* ``make_extra_regressors``-- ``J_regressor_extra`` (9x6890), ``J_regressor_h36m`` (17x6890)
                              (reference config.py:14-15)
* ``make_mean_params``     -- ``smpl_mean_params.npz`` content (reference model/hmr.py:100-106)
* ``make_basemodel``       -- a ``{'model': {'module.<name>': tensor}}`` checkpoint
                              (reference base_adaptor.py:116-121)
* ``make_exemplar_bank`` / ``make_clusters`` -- the H36M retrieval store
                              (reference base_adaptor.py:74-96, :450-555)
* ``SyntheticStream``      -- 3DPW-shaped frame dicts (reference boa_dataset/pw3d.py:115-125)

This module is data generation only: the label generator below poses the body
with a small numpy LBS so that keypoints are consistent with (pose, betas); it is
never on the adaptation path, which runs through the CUDA library exclusively.
"""
from collections import OrderedDict
import math

import numpy as np
import torch

from . import constants, layout

SEED = 22                      # reference dynaboa_benchmark.py:20
NV = constants.NUM_VERTS

# Approximate SMPL rest-pose joint locations (metres, y up), used to shape the stand-in body.
_REST_JOINTS = np.array([
    [0.00, 0.00, 0.00], [0.07, -0.09, 0.00], [-0.07, -0.09, 0.00], [0.00, 0.11, -0.01],
    [0.10, -0.47, 0.01], [-0.10, -0.47, 0.01], [0.00, 0.25, 0.00], [0.09, -0.87, -0.03],
    [-0.09, -0.87, -0.03], [0.00, 0.30, 0.02], [0.11, -0.93, 0.09], [-0.11, -0.93, 0.09],
    [0.00, 0.51, -0.01], [0.08, 0.42, 0.00], [-0.08, 0.42, 0.00], [0.00, 0.60, 0.04],
    [0.18, 0.45, -0.01], [-0.18, 0.45, -0.01], [0.44, 0.44, -0.02], [-0.44, 0.44, -0.02],
    [0.69, 0.44, -0.02], [-0.69, 0.44, -0.02], [0.77, 0.43, -0.02], [-0.77, 0.43, -0.02],
], dtype=np.float64)

_GENDER_SEED = {'neutral': 0, 'male': 1, 'female': 2}


def _sparse_rows(rng, nrows, anchors, verts, k):
    """Row-stochastic (nrows, NV) matrix; row r mixes the k vertices nearest to anchors[r]."""
    out = np.zeros((nrows, NV), dtype=np.float64)
    for r in range(nrows):
        d = np.linalg.norm(verts - anchors[r], axis=1)
        idx = np.argsort(d)[:k]
        w = rng.random(k) + 0.05
        out[r, idx] = w / w.sum()
    return out.astype(np.float32)


def make_smpl_model(gender='neutral', seed=SEED):
    """SMPL-shaped model buffers as float32/int64 numpy arrays."""
    rng = np.random.default_rng(seed * 1000 + 17 * _GENDER_SEED[gender] + 1)
    scale = {'neutral': 1.0, 'male': 1.04, 'female': 0.95}[gender]
    rest = _REST_JOINTS * scale
    owner = rng.permutation(NV) % 24
    v_template = rest[owner] + rng.normal(0.0, 0.045, size=(NV, 3))
    # skinning weights: gaussian falloff to the joints, top-4, renormalised
    d2 = ((v_template[:, None, :] - rest[None, :, :]) ** 2).sum(-1)
    w = np.exp(-d2 / (2 * 0.09 ** 2)) + 1e-12
    drop = np.argsort(-w, axis=1)[:, 4:]
    np.put_along_axis(w, drop, 0.0, axis=1)
    lbs_weights = w / w.sum(1, keepdims=True)
    J_regressor = _sparse_rows(rng, 24, rest, v_template, 160)
    shapedirs = rng.normal(0.0, 0.012, size=(NV, 3, 10))
    posedirs = rng.normal(0.0, 0.004, size=(207, NV * 3))
    faces = rng.integers(0, NV, size=(13776, 3), dtype=np.int64)
    return OrderedDict(
        v_template=v_template.astype(np.float32),
        shapedirs=shapedirs.astype(np.float32),
        posedirs=posedirs.astype(np.float32),
        J_regressor=J_regressor,
        parents=np.asarray(constants.SMPL_PARENTS, dtype=np.int64),
        lbs_weights=lbs_weights.astype(np.float32),
        faces=faces,
    )


def make_extra_regressors(seed=SEED):
    rng = np.random.default_rng(seed * 1000 + 101)
    body = make_smpl_model('neutral', seed)
    v = body['v_template'].astype(np.float64)
    rest = _REST_JOINTS
    # 9 extra joints (hips, LSP neck / head top, pelvis, thorax, spine, jaw, head)
    extra_anchor = np.stack([rest[2] * [1.3, 1, 1], rest[1] * [1.3, 1, 1], rest[12], rest[15] + [0, 0.1, 0],
                             rest[0], rest[9] + [0, 0.08, 0], rest[3], rest[15] + [0, -0.03, 0.06],
                             rest[15] + [0, 0.04, 0]])
    # 17 H36M joints: pelvis, legs, spine, head, arms
    h36m_ids = [0, 2, 5, 8, 1, 4, 7, 6, 12, 15, 15, 16, 18, 20, 17, 19, 21]
    h36m_anchor = rest[h36m_ids] + rng.normal(0, 0.01, size=(17, 3))
    return OrderedDict(
        J_regressor_extra=_sparse_rows(rng, 9, extra_anchor, v, 220),
        J_regressor_h36m=_sparse_rows(rng, 17, h36m_anchor, v, 220),
    )


def make_mean_params():
    return OrderedDict(
        pose=np.tile(np.array([1, 0, 0, 1, 0, 0], dtype=np.float32), 24),
        shape=np.zeros(10, dtype=np.float32),
        cam=np.array([0.9, 0.0, 0.0], dtype=np.float32),
    )


def make_basemodel(seed=SEED, dec_gain=0.2, prefix='module.'):
    """Random stand-in for ``data/basemodel.pt``: conv ~ N(0, 2/(k*k*cout)) as the reference
    init (model/hmr.py:92-95), non-trivial GroupNorm affine, default Linear init for fc1/fc2
    and xavier-uniform decoders (gain raised from 0.01 so predicted poses are O(0.1 rad))."""
    g = torch.Generator().manual_seed(seed * 1000 + 7)
    sd = OrderedDict()
    for name, shp in layout.param_shapes().items():
        if len(shp) == 4:
            cout, _, k, _ = shp
            t = torch.randn(shp, generator=g) * math.sqrt(2.0 / (k * k * cout))
        elif 'bn' in name or 'downsample.1' in name:
            t = torch.randn(shp, generator=g) * 0.1
            if name.endswith('weight'):
                t = t + 1.0
        elif name.split('.')[0] in ('fc1', 'fc2'):
            fan_in = layout.HEAD_IN if name.startswith('fc1') else layout.HEAD_HID
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        else:  # decoders
            if len(shp) == 2:
                bound = dec_gain * math.sqrt(6.0 / (shp[0] + shp[1]))
            else:
                bound = 1.0 / math.sqrt(layout.HEAD_HID)
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        sd[prefix + name] = t.float().contiguous()
    mp = make_mean_params()
    sd[prefix + 'init_pose'] = torch.from_numpy(mp['pose']).unsqueeze(0)
    sd[prefix + 'init_shape'] = torch.from_numpy(mp['shape']).unsqueeze(0)
    sd[prefix + 'init_cam'] = torch.from_numpy(mp['cam']).unsqueeze(0)
    return {'model': sd}


# --------------------------------------------------------------------------------------
# label generation (numpy LBS; data generation only)
# --------------------------------------------------------------------------------------
def _rodrigues_np(aa):
    """smplx-style Rodrigues: (N,3) -> (N,3,3)."""
    angle = np.linalg.norm(aa + 1e-8, axis=1, keepdims=True)
    k = aa / angle
    K = np.zeros((aa.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(angle)[:, :, None], np.cos(angle)[:, :, None]
    return np.eye(3)[None] + s * K + (1 - c) * (K @ K)


def _datagen_joints49(body, extra, pose72, betas):
    """49 SPIN joints of a posed body, float64 numpy.  Used only to synthesise labels."""
    vt = body['v_template'].astype(np.float64)
    v_shaped = vt + np.einsum('l,mkl->mk', betas, body['shapedirs'].astype(np.float64))
    J = body['J_regressor'].astype(np.float64) @ v_shaped
    R = _rodrigues_np(pose72.reshape(24, 3).astype(np.float64))
    pf = (R[1:] - np.eye(3)[None]).reshape(-1)
    v_posed = v_shaped + (pf @ body['posedirs'].astype(np.float64)).reshape(NV, 3)
    parents = body['parents']
    Gr, Gt = [None] * 24, [None] * 24
    Gr[0], Gt[0] = R[0], J[0]
    for j in range(1, 24):
        p = parents[j]
        Gr[j] = Gr[p] @ R[j]
        Gt[j] = Gr[p] @ (J[j] - J[p]) + Gt[p]
    Gr, Gt = np.stack(Gr), np.stack(Gt)
    At = Gt - np.einsum('jab,jb->ja', Gr, J)
    W = body['lbs_weights'].astype(np.float64)
    Tr = np.einsum('vj,jab->vab', W, Gr)
    Tt = W @ At
    verts = np.einsum('vab,vb->va', Tr, v_posed) + Tt
    j45 = np.concatenate([Gt, verts[constants.SMPL_EXTRA_VERTEX_IDS]], 0)
    j54 = np.concatenate([j45, extra['J_regressor_extra'].astype(np.float64) @ verts], 0)
    return j54[constants.JOINT_MAP_49], verts


def _project_norm(j3d, cam):
    """Weak-perspective camera of reference base_adaptor.py:160-170 -> coords in [-1,1]."""
    t = np.array([cam[1], cam[2], 2 * constants.FOCAL_LENGTH / (constants.IMG_RES * cam[0] + 1e-9)])
    p = j3d + t
    return constants.FOCAL_LENGTH * p[:, :2] / p[:, 2:3] / (constants.IMG_RES / 2.0)


def _quantise_kp(p_norm):
    """Reference keypoints are integer pixels mapped to [-1,1] (boa_dataset/pw3d.py:157-160)."""
    pix = np.clip(np.rint((p_norm + 1.0) * constants.IMG_RES / 2.0), 1, constants.IMG_RES)
    return 2.0 * pix / constants.IMG_RES - 1.0


def _norm_image(rng, shape):
    img = rng.random(shape, dtype=np.float32)
    mean = np.asarray(constants.IMG_NORM_MEAN, np.float32).reshape(3, 1, 1)
    std = np.asarray(constants.IMG_NORM_STD, np.float32).reshape(3, 1, 1)
    return (img - mean) / std


def make_exemplar_bank(n=64, seed=SEED):
    """Pre-decoded stand-in for the H36M ``SourceDataset`` (reference base_adaptor.py:450-555):
    tensors with a leading item axis; ``retrieval`` gathers rows instead of reading JPEGs."""
    rng = np.random.default_rng(seed * 1000 + 303)
    body, extra = make_smpl_model('neutral', seed), make_extra_regressors(seed)
    img = np.stack([_norm_image(rng, (3, 224, 224)) for _ in range(n)])
    pose = rng.normal(0, 0.2, size=(n, 72))
    betas = rng.normal(0, 0.6, size=(n, 10))
    kps = np.zeros((n, 49, 3), np.float32)
    s3d = np.zeros((n, 24, 4), np.float32)
    for i in range(n):
        j49, _ = _datagen_joints49(body, extra, pose[i], betas[i])
        cam = np.array([0.85 + 0.1 * rng.random(), rng.normal(0, 0.03), rng.normal(0, 0.03)])
        kps[i, 25:, :2] = _quantise_kp(_project_norm(j49[25:], cam))
        kps[i, 25:, 2] = 1.0
        s3d[i, :, :3] = j49[25:]
        s3d[i, :, 3] = 1.0
    # a few invisible joints so the confidence masks are exercised
    hide = rng.random((n, 24)) < 0.08
    kps[:, 25:, 2][hide] = 0.0
    return OrderedDict(
        img=torch.from_numpy(img.astype(np.float32)),
        pose_3d=torch.from_numpy(s3d),
        betas=torch.from_numpy(betas.astype(np.float32)),
        pose=torch.from_numpy(pose.astype(np.float32)),
        keypoints=torch.from_numpy(kps),
    )


def make_clusters(n_items=64, k=6, seed=SEED):
    """Cluster centres (K,2048) and per-cluster item lists (reference base_adaptor.py:74-80)."""
    rng = np.random.default_rng(seed * 1000 + 404)
    centers = np.abs(rng.normal(0.5, 0.3, size=(k, 2048))).astype(np.float32)
    perm = rng.permutation(n_items)
    index = [sorted(int(x) for x in perm[c::k]) for c in range(k)]
    return OrderedDict(centers=centers, index=index)


class SyntheticStream:
    """Temporally smooth 3DPW-shaped frame stream; ``stream[t]`` -> batch dict with a leading
    batch axis of ``batch_size`` (independent sequences per batch row)."""

    def __init__(self, length=16, batch_size=1, seed=SEED, rank=0):
        self.length, self.batch_size = length, batch_size
        rng = np.random.default_rng(seed * 1000 + 505 + 7919 * rank)
        bodies = {g: make_smpl_model(g, seed) for g in ('male', 'female')}
        extra = make_extra_regressors(seed)
        B = batch_size
        self.gender = rng.integers(0, 2, size=B).astype(np.int32)
        self.betas = rng.normal(0, 0.6, size=(B, 10))
        pose = rng.normal(0, 0.2, size=(B, 72))
        cam = np.stack([0.85 + 0.1 * rng.random(B), rng.normal(0, 0.03, B), rng.normal(0, 0.03, B)], 1)
        img = np.stack([_norm_image(rng, (3, 224, 224)) for _ in range(B)])
        self.frames = []
        for t in range(length):
            if t > 0:
                img = 0.9 * img + 0.1 * np.stack([_norm_image(rng, (3, 224, 224)) for _ in range(B)])
                pose = pose + rng.normal(0, 0.02, size=pose.shape)
            j2d = np.zeros((B, 49, 3), np.float32)
            for b in range(B):
                body = bodies['male' if self.gender[b] == 0 else 'female']
                j49, _ = _datagen_joints49(body, extra, pose[b], self.betas[b])
                j2d[b, :, :2] = _quantise_kp(_project_norm(j49, cam[b]))
                j2d[b, :, 2] = 1.0
            op = j2d.copy()
            op[:, :, :2] += rng.normal(0, 0.01, size=(B, 49, 2)).astype(np.float32)
            self.frames.append(OrderedDict(
                image=torch.from_numpy(img.astype(np.float32).copy()),
                smpl_j2d=torch.from_numpy(j2d),
                op_j2d=torch.from_numpy(op),
                pose=torch.from_numpy(pose.astype(np.float32).copy()),
                betas=torch.from_numpy(self.betas.astype(np.float32).copy()),
                gender=torch.from_numpy(self.gender.copy()),
                j3d=torch.zeros(B, 24, 4),
                bbox=torch.tensor([[112.0, 112.0, 200.0]]).repeat(B, 1),
                imgname=[f'synthetic/seq_{rank}_{b}/frame_{t:05d}.jpg' for b in range(B)],
                dataset_name=['3dpw'] * B,
            ))

    def __len__(self):
        return self.length

    def __getitem__(self, t):
        return self.frames[t]

    def __iter__(self):
        return iter(self.frames)


def write_asset_dir(root, seed=SEED, n_exemplars=64, k=6):
    """Materialise every asset file ``BaseAdaptor.__init__`` loads (paths in ``config``)."""
    import os
    os.makedirs(os.path.join(root, 'smpl'), exist_ok=True)
    os.makedirs(os.path.join(root, 'retrieval_res'), exist_ok=True)
    for g in ('neutral', 'male', 'female'):
        np.savez(os.path.join(root, 'smpl', f'SMPL_{g.upper()}.npz'), **make_smpl_model(g, seed))
    ex = make_extra_regressors(seed)
    np.save(os.path.join(root, 'J_regressor_extra.npy'), ex['J_regressor_extra'])
    np.save(os.path.join(root, 'J_regressor_h36m.npy'), ex['J_regressor_h36m'])
    np.savez(os.path.join(root, 'smpl_mean_params.npz'), **make_mean_params())
    torch.save(make_basemodel(seed), os.path.join(root, 'basemodel.pt'))
    torch.save(dict(make_exemplar_bank(n_exemplars, seed)), os.path.join(root, 'retrieval_res', 'exemplar_bank.pt'))
    cl = make_clusters(n_exemplars, k, seed)
    torch.save({'centers': cl['centers'], 'index': cl['index']},
               os.path.join(root, 'retrieval_res', 'clusters.pt'))
    return root
