"""Oracle restatement of the HMR regressor forward (test infrastructure only).

Functional form of reference model/hmr.py: ``Bottleneck.forward`` (:40-60), ``HMR.forward``
(:127-181).  Parameters come as a flat dict ``name -> tensor`` with the reference's state_dict
names (no ``module.`` prefix), which is what lets the same function serve fast weights in the
MAML restatement.  Dropout is expressed through explicit keep-masks (already scaled by 1/(1-p))
so the stochastic teacher of the benchmark driver can be reproduced.
"""
import torch
import torch.nn.functional as F

from . import geometry_ref

GROUPS, EPS = 4, 1e-5          # model/hmr.py:18  nn.GroupNorm(32 // 8, planes)
BLOCKS = (3, 4, 6, 3)          # model/hmr.py:322


def _gn(x, p, name):
    return F.group_norm(x, GROUPS, p[name + '.weight'], p[name + '.bias'], EPS)


def _bottleneck(x, p, pre, stride, has_ds):
    out = F.relu(_gn(F.conv2d(x, p[pre + '.conv1.weight']), p, pre + '.bn1'))
    out = F.relu(_gn(F.conv2d(out, p[pre + '.conv2.weight'], stride=stride, padding=1), p, pre + '.bn2'))
    out = _gn(F.conv2d(out, p[pre + '.conv3.weight']), p, pre + '.bn3')
    res = x
    if has_ds:
        res = _gn(F.conv2d(x, p[pre + '.downsample.0.weight'], stride=stride), p, pre + '.downsample.1')
    return F.relu(out + res)


def backbone(x, p):
    """model/hmr.py:138-156.  Returns (xf (B,2048), [stem_conv, layer1..4 outputs])."""
    feats = []
    y = F.conv2d(x, p['conv1.weight'], stride=2, padding=3)
    feats.append(y)
    y = F.max_pool2d(F.relu(_gn(y, p, 'bn1')), kernel_size=3, stride=2, padding=1)
    for li, nblk in enumerate(BLOCKS):
        for bi in range(nblk):
            stride = 2 if (li > 0 and bi == 0) else 1
            y = _bottleneck(y, p, f'layer{li + 1}.{bi}', stride, bi == 0)
        feats.append(y)
    xf = F.avg_pool2d(y, 7, stride=1).flatten(1)
    return xf, feats


def regressor(xf, p, init_pose, init_shape, init_cam, n_iter=3, masks=None):
    """model/hmr.py:158-172.  ``masks``: None (eval) or list of n_iter (m1, m2) scaled keep-masks."""
    pose, shape, cam = init_pose, init_shape, init_cam
    feats = []
    for i in range(n_iter):
        xc = torch.cat([xf, pose, shape, cam], 1)
        xc = F.linear(xc, p['fc1.weight'], p['fc1.bias'])
        feats.append(xc.clone())
        if masks is not None:
            xc = xc * masks[i][0]
        feats.append(xc.clone())
        xc = F.linear(xc, p['fc2.weight'], p['fc2.bias'])
        feats.append(xc.clone())
        if masks is not None:
            xc = xc * masks[i][1]
        pose = F.linear(xc, p['decpose.weight'], p['decpose.bias']) + pose
        shape = F.linear(xc, p['decshape.weight'], p['decshape.bias']) + shape
        cam = F.linear(xc, p['deccam.weight'], p['deccam.bias']) + cam
    return pose, shape, cam, feats


def forward(x, p, need_feature=False, n_iter=3, masks=None, return_pose6d=False):
    """model/hmr.py:127-181.  ``p`` must also hold the init_pose/init_shape/init_cam buffers."""
    B = x.shape[0]
    xf, feats = backbone(x, p)
    feats.append(xf)
    pose, shape, cam, hfeats = regressor(xf, p, p['init_pose'].expand(B, -1), p['init_shape'].expand(B, -1),
                                         p['init_cam'].expand(B, -1), n_iter, masks)
    rotmat = geometry_ref.rot6d_to_rotmat(pose).view(B, 24, 3, 3)
    out = (rotmat, shape, cam)
    if need_feature:
        out = out + (feats + hfeats,)
    if return_pose6d:
        out = out + (pose,)
    return out


def strip_prefix(sd, prefix='module.'):
    return {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in sd.items()}
