"""Exemplar store of the retrieval path (N4): the H36M annotation file of the reference turned ONCE into the device-resident
bank ``retrieval`` gathers rows from (reference: ``SourceDataset.__getitem__`` base_adaptor.py:450-555 decodes a JPEG, crops and
resizes it on the CPU for every retrieved exemplar of every frame; ``load_h36_cluster_res`` :70-80).

    python -m dynaboa_b200.exemplars data/retrieval_res/h36m_random_sample_center_10_10.pt data/h36m data/retrieval_res/exemplar_bank.pt \\
        [data/retrieval_res/cluster_res_random_sample_center_10_10_potocol2.pt]

Per item, exactly what ``SourceDataset.__getitem__`` returns to the adaptation path -- ``img`` (crop / resize / normalise,
on the GPU: dynaboa_b200.dataprocess), ``keypoints`` (25 zero OpenPose rows + the 24 ground-truth joints, transformed to crop
coordinates), ``pose`` (72), ``betas`` (10), ``pose_3d`` (24, 4) -- stacked along a leading item axis in fp32 (602 KB per
exemplar: 180 GB of HBM hold ~300k crops).  With the optional cluster file only the items its index refers to are kept and
the index is rewritten over bank rows (``<out>_clusters.pt``).

Note: the reference's ``SourceDataset.read_image`` raises on every call (``if not img:`` on an array, :513-519); the reading
rule restated here is the evident intent (BGR -> RGB float32), the one pw3d.py:81-85 uses.
"""
import os
import sys

import numpy as np
import torch

from . import dataprocess


def read_image(path):
    import cv2
    img = cv2.imread(path)
    if img is None:
        raise FileNotFoundError(path)
    return img[:, :, ::-1].copy()                      # uint8 RGB; the crop kernel converts


def load_annotations(path):
    """The reference's joblib pickle (``imgname, scale, center, pose, shape, S, part[, gender]``) or an npz of the same keys."""
    if path.endswith('.npz'):
        return dict(np.load(path, allow_pickle=True))
    import joblib
    return joblib.load(path)


def build_bank(annotations, img_dir, indices=None, device='cuda', reader=read_image):
    """dict of stacked tensors (on ``device``), rows in the order of ``indices`` (default: every item)."""
    a = annotations
    n = len(a['imgname'])
    indices = list(range(n)) if indices is None else [int(i) for i in indices]
    dev = torch.device(device)
    out = {k: [] for k in ('img', 'keypoints', 'pose', 'betas', 'pose_3d')}
    for i in indices:
        center, scale = [float(c) for c in a['center'][i]], float(a['scale'][i])
        frame = torch.from_numpy(reader(os.path.join(img_dir, str(a['imgname'][i])))).to(dev)
        kp = np.concatenate([np.zeros((25, 3)), np.asarray(a['part'][i], dtype=np.float64)], 0)       # :466-468
        out['img'].append(dataprocess.crop(frame, center, scale))
        out['keypoints'].append(dataprocess.j2d_processing(torch.from_numpy(kp).float().to(dev), center, scale))
        out['pose'].append(torch.from_numpy(np.asarray(a['pose'][i], dtype=np.float64).astype(np.float32)).to(dev))   # rot = 0: unchanged
        out['betas'].append(torch.from_numpy(np.asarray(a['shape'][i], dtype=np.float64).astype(np.float32)).to(dev))
        out['pose_3d'].append(torch.from_numpy(np.asarray(a['S'][i]).astype(np.float32)).to(dev))
    return {k: torch.stack(v).contiguous() for k, v in out.items()}


def remap_clusters(clusters, indices):
    """Cluster index over annotation ids -> over bank rows (``indices[row] = annotation id``)."""
    row = {int(a): r for r, a in enumerate(indices)}
    return {'centers': clusters['centers'], 'index': [[row[int(i)] for i in ids if int(i) in row] for ids in clusters['index']]}


def main(argv):
    if len(argv) < 3:
        print(__doc__)
        return 2
    annot, img_dir, out = argv[:3]
    a = load_annotations(annot)
    indices = None
    if len(argv) > 3:                                  # cluster file: keep only the items it refers to
        import joblib
        cl = joblib.load(argv[3])
        indices = sorted({int(i) for ids in cl['index'] for i in ids})
    bank = build_bank(a, img_dir, indices)
    torch.save({k: v.cpu() for k, v in bank.items()}, out)
    if indices is not None:
        torch.save(remap_clusters(cl, indices), os.path.splitext(out)[0] + '_clusters.pt')
    print(f'{out}: {bank["img"].shape[0]} exemplars, {sum(v.numel() * 4 for v in bank.values()) / 2**20:.1f} MiB')
    return 0


if __name__ == '__main__':
    sys.exit(main(sys.argv[1:]))
