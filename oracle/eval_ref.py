"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  CPU restatement of the evaluation arithmetic of reference
dynaboa_benchmark.py:217-240 (Adaptor.inference) and utils/pose_utils.py:9-64 (similarity Procrustes):
H36M-regressor joints of the predicted and ground-truth meshes, pelvis centring, 14-joint selection, MPJPE,
PA-MPJPE and PVE.  Pinned against the reference's own ``compute_similarity_transform_batch`` by
oracle/make_golden.py (fixture tests/golden/eval_metrics.npz)."""
import numpy as np


def similarity_transform(S1, S2):
    """utils/pose_utils.py:9-56 on (N,3) arrays: S1 mapped closest to S2 by scale, rotation, translation."""
    A, B = S1.T, S2.T
    muA, muB = A.mean(axis=1, keepdims=True), B.mean(axis=1, keepdims=True)
    A0, B0 = A - muA, B - muB
    var = np.sum(A0 ** 2)
    K = A0.dot(B0.T)
    U, _, Vh = np.linalg.svd(K)
    V = Vh.T
    Z = np.eye(3)
    Z[-1, -1] *= np.sign(np.linalg.det(U.dot(V.T)))
    R = V.dot(Z.dot(U.T))
    scale = np.trace(R.dot(K)) / var
    t = muB - scale * R.dot(muA)
    return (scale * R.dot(A) + t).T


def eval_metrics(pred_verts, gt_verts_joints, gt_verts_pve, J_regressor, joint_map):
    """:217-240 -- returns (mpjpe[B], pampjpe[B], pve scalar) in the unit of the meshes."""
    gt_k = np.matmul(J_regressor[None], gt_verts_joints)
    gt_k = gt_k[:, joint_map] - gt_k[:, [0]]
    pr_k = np.matmul(J_regressor[None], pred_verts)
    pr_k = pr_k[:, joint_map] - pr_k[:, [0]]
    mpjpe = np.sqrt(((pr_k - gt_k) ** 2).sum(-1)).mean(-1)
    hat = np.stack([similarity_transform(pr_k[i], gt_k[i]) for i in range(pr_k.shape[0])])
    pampjpe = np.sqrt(((hat - gt_k) ** 2).sum(-1)).mean(-1)
    pve = np.sqrt(((gt_verts_pve - pred_verts) ** 2).sum(2)).mean()
    return mpjpe, pampjpe, pve
