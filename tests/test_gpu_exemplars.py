"""Exemplar store (N4): the reference's H36M annotation layout ingested into the device-resident bank on the GPU input side,
against the CPU restatement of ``SourceDataset.__getitem__`` (reference base_adaptor.py:450-555) item by item."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def make_annotations(root, n=5):
    import cv2
    rng = np.random.default_rng(21)
    os.makedirs(os.path.join(root, 'S1'), exist_ok=True)
    a = dict(imgname=[], center=[], scale=[], pose=[], shape=[], S=[], part=[])
    frames = []
    for i in range(n):
        frame = cv2.GaussianBlur(rng.uniform(0, 255, size=(260, 300, 3)).astype(np.float32), (0, 0), 2).astype(np.uint8)
        name = f'S1/img_{i:04d}.png'
        cv2.imwrite(os.path.join(root, name), frame[:, :, ::-1])
        center = np.array([150.0 + 5 * i, 130.0 - 3 * i])
        a['imgname'].append(name); a['center'].append(center); a['scale'].append(0.85 + 0.05 * i)
        a['pose'].append(rng.normal(0, 0.3, 72)); a['shape'].append(rng.normal(0, 0.5, 10))
        a['S'].append(np.concatenate([rng.normal(0, 0.4, (24, 3)), np.ones((24, 1))], 1))
        a['part'].append(np.concatenate([center + rng.uniform(-80, 80, (24, 2)), (rng.random((24, 1)) > 0.2).astype(np.float64)], 1))
        frames.append(frame)
    return {k: np.array(v) for k, v in a.items()}, frames


def test_bank_matches_the_source_dataset_items(tmp_path):
    from dynaboa_b200 import exemplars
    from oracle import dataprocess_ref as R
    a, frames = make_annotations(str(tmp_path))
    path = str(tmp_path / 'annot.npz')
    np.savez(path, **a)
    idx = [3, 0, 4]
    bank = exemplars.build_bank(exemplars.load_annotations(path), str(tmp_path), idx)
    assert bank['img'].shape == (3, 3, 224, 224) and bank['keypoints'].shape == (3, 49, 3) and bank['pose_3d'].shape == (3, 24, 4)
    for row, i in enumerate(idx):
        center, scale = list(a['center'][i]), float(a['scale'][i])
        img = R.rgb_processing(frames[i].astype(np.float32), center, scale)        # crop + /255 + Normalize, as __getitem__ :493-499
        assert rel_err(bank['img'][row], img) < 1e-5, i
        kp = R.j2d_processing(np.concatenate([np.zeros((25, 3)), a['part'][i]], 0), center, scale)
        assert np.array_equal(bank['keypoints'][row].cpu().numpy(), kp.astype(np.float32)), i
        assert np.array_equal(bank['pose'][row].cpu().numpy(), a['pose'][i].astype(np.float32))
        assert np.array_equal(bank['betas'][row].cpu().numpy(), a['shape'][i].astype(np.float32))
        assert np.array_equal(bank['pose_3d'][row].cpu().numpy(), a['S'][i].astype(np.float32))
    cl = exemplars.remap_clusters({'centers': np.zeros((2, 2048), np.float32), 'index': [[0, 3, 1], [4]]}, idx)
    assert cl['index'] == [[1, 0], [2]]
