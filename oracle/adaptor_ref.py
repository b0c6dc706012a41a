"""Oracle restatement of the bilevel adaptation loop (test infrastructure only).

CPU torch restatement of reference base_adaptor.py (``BaseAdaptor``: projection :160-170,
history :173-180, decode_smpl_params :183-190, update_teacher :193-201, cal_feature_diff :211-219,
lower/upper_level_adaptation :222-317, cal_teacher_loss :320-343, adapt_on_labeled_data :346-376,
cal_motion_loss :379-398, priors :401-409, cal_s3d_loss :412-422, retrieval :82-96) and of the
driver's ``Adaptor.adaptation`` / ``Adaptor.inference`` (reference dynaboa_benchmark.py:126-201,
:204-262), on top of the hmr / smplx / learn2learn restatements in this package.

This is also the timed CPU baseline ("port") of ``bench.py``.  It is never imported by the product.
"""
import random
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from . import geometry_ref as G
from . import hmr_ref, l2l_ref, prior_ref, smplx_ref


def default_options(**over):
    """Flag defaults of reference dynaboa_benchmark.py:16-65."""
    o = dict(seed=22, batch_size=1, lr=3e-6, beta1=0.5, beta2=0.9, use_boa=1, fastlr=8e-6, inner_step=1,
             s2dloss_weight=10.0, shape_prior_weight=2e-6, pose_prior_weight=1e-4,
             use_frame_losses_lower=1, use_frame_losses_upper=1, use_temporal_losses_lower=0,
             use_temporal_losses_upper=1, sample_num=1, retrieval=1, dynamic_boa=1,
             cos_sim_threshold=3.1e-4, optim_steps=7, lower_level_mixtrain=1, upper_level_mixtrain=1,
             labelloss_weight=0.1, use_meanteacher=1, alpha=0.1, teacherloss_weight=0.1, use_motion=1,
             interval=5, motionloss_weight=0.8, teacher_dropout=1)
    o.update(over)
    return SimpleNamespace(**o)


def similarity_transform(S1, S2):
    """reference utils/pose_utils.py:9-57 for (N,3) point sets (float numpy)."""
    X1, X2 = S1.T, S2.T
    mu1, mu2 = X1.mean(1, keepdims=True), X2.mean(1, keepdims=True)
    Y1, Y2 = X1 - mu1, X2 - mu2
    var1 = (Y1 ** 2).sum()
    K = Y1 @ Y2.T
    U, _, Vh = np.linalg.svd(K)
    V = Vh.T
    Z = np.eye(3)
    Z[-1, -1] *= np.sign(np.linalg.det(U @ V.T))
    R = V @ Z @ U.T
    scale = np.trace(R @ K) / var1
    t = mu2 - scale * (R @ mu1)
    return (scale * (R @ X1) + t).T


class OracleAdaptor:
    def __init__(self, options, checkpoint, smpl_models, regressors, gmm, bank=None, clusters=None,
                 dtype=torch.float32, joint_map=None, vertex_ids=None, h36m_to_j14=None):
        self.o = options
        self.dtype = dtype
        random.seed(options.seed)                                  # base_adaptor.py:98-109
        sd = hmr_ref.strip_prefix(checkpoint['model'])
        self.buffers = {k: sd[k].to(dtype) for k in ('init_pose', 'init_shape', 'init_cam')}
        self.theta = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items() if k not in self.buffers}
        self.teacher = {k: v.detach().clone() for k, v in self.theta.items()}  # set_teacher :151-158
        self.optimizer = torch.optim.Adam(list(self.theta.values()), lr=options.lr,
                                          betas=(options.beta1, options.beta2))   # :126
        t = lambda a: torch.as_tensor(a, dtype=dtype)
        self.smpl = {}
        for g, m in smpl_models.items():
            self.smpl[g] = {k: (torch.as_tensor(v, dtype=torch.long) if k == 'parents' else t(v))
                            for k, v in m.items() if k != 'faces'}
        self.J_extra = t(regressors['J_regressor_extra'])
        self.J_h36m = t(regressors['J_regressor_h36m'])
        self.joint_map = torch.as_tensor(joint_map, dtype=torch.long)
        self.vertex_ids = torch.as_tensor(vertex_ids, dtype=torch.long)
        self.h36m_to_j14 = list(h36m_to_j14)
        self.gmm = prior_ref.gmm_constants(gmm, dtype)
        self.bank = None if bank is None else {k: v.to(dtype) for k, v in bank.items()}
        if clusters is not None:
            self.centers = t(clusters['centers'])
            self.index = clusters['index']
        self.history = {}
        self.global_step = 0
        self.fit_losses = {}
        self.kp2dlosses_lower, self.kp2dlosses_upper = [], {}
        self.mask_fn = None          # callable(B) -> [(m1, m2)] * 3 scaled keep-masks for the teacher
        # data-parallel emulation (oracle/dp_ref.py): called between loss.backward() and optimizer.step() / with the per-feature
        # (a.b, |a|^2, |b|^2) sums of the dynamic test; None = single stream (the reference's semantics)
        self.grad_hook = None
        self.cos_hook = None
        self.mask_gen = torch.Generator().manual_seed(options.seed)

    # ------------------------------------------------------------------ model pieces
    def with_buffers(self, params):
        p = dict(params)
        p.update(self.buffers)
        return p

    def teacher_masks(self, B):
        if not self.o.teacher_dropout:
            return None
        if self.mask_fn is not None:
            return self.mask_fn(B)
        out = []
        for _ in range(3):
            out.append(tuple((torch.rand(B, 1024, generator=self.mask_gen) >= 0.5).to(self.dtype) * 2.0
                             for _ in range(2)))
        return out

    def decode(self, rotmat, beta, gender='neutral', pose2rot=False):
        """base_adaptor.py:183-190."""
        if pose2rot:
            out = smplx_ref.smpl_forward(self.smpl[gender], self.J_extra, self.joint_map, self.vertex_ids,
                                         beta, rotmat[:, 3:], rotmat[:, :3], True)
        else:
            out = smplx_ref.smpl_forward(self.smpl[gender], self.J_extra, self.joint_map, self.vertex_ids,
                                         beta, rotmat[:, 1:], rotmat[:, 0].unsqueeze(1), False)
        return out.joints, out.vertices

    def project(self, cam, s3d):
        return G.weak_perspective_project(cam, s3d)[1]

    # ------------------------------------------------------------------ losses
    def shape_prior(self, betas):
        return (betas ** 2).sum(-1).mean()                                          # :401-402

    def pose_prior(self, rotmat):
        aa = G.rotation_matrix_to_angle_axis(rotmat[:, 1:].contiguous().view(-1, 3, 3)).view(-1, 69)
        return prior_ref.merged_nll(aa, self.gmm).mean()                            # :405-409

    @staticmethod
    def s3d_loss(pred, gt, conf):
        gt = gt - ((gt[:, 2] + gt[:, 3]) / 2)[:, None, :]                           # :412-422
        pred = pred - ((pred[:, 2] + pred[:, 3]) / 2)[:, None, :]
        return (conf * (pred - gt) ** 2).mean()

    def frame_losses(self, s2d, kp, shape, rotmat, tag):
        conf = kp[:, 25:, -1].unsqueeze(-1)
        s2dloss = (((s2d[:, 25:] - kp[:, 25:, :-1]) ** 2) * conf).mean()            # :234 / :283
        sp, pp = self.shape_prior(shape), self.pose_prior(rotmat)
        loss = s2dloss * self.o.s2dloss_weight + sp * self.o.shape_prior_weight + pp * self.o.pose_prior_weight
        self.fit_losses[f'{tag}/s2dloss'] = s2dloss
        self.fit_losses[f'{tag}/shape_prior'] = sp
        self.fit_losses[f'{tag}/pose_prior'] = pp
        self.fit_losses[f'{tag}/unlabelloss'] = loss
        return loss, s2dloss

    def teacher_loss(self, image, rotmat, shape, s2d, s3d):
        """:320-343 -- the teacher is never put in eval mode by the benchmark driver, so its
        dropout layers are live."""
        with torch.no_grad():
            t_rot, t_shape, t_cam = hmr_ref.forward(image, self.with_buffers(self.teacher),
                                                    masks=self.teacher_masks(image.shape[0]))
            t_s3d, _ = self.decode(t_rot, t_shape)
            t_s2d = self.project(t_cam, t_s3d)
        l2d, l3d = F.mse_loss(s2d, t_s2d), F.mse_loss(t_s3d, s3d)
        lsh, lpo = F.mse_loss(shape, t_shape), F.mse_loss(rotmat, t_rot)
        loss = l2d * 5 + l3d * 5 + lsh * 0.001 + lpo * 1
        self.fit_losses.update({'teacher/s2dloss': l2d, 'teacher/s3dloss': l3d, 'teacher/shape_loss': lsh,
                                'teacher/pose_loss': lpo, 'teacher/loss': loss})
        return loss

    def motion_loss(self, params, s2d_gt_part, kp_part):
        """:379-398 -- forward on frame t-interval with the same (fast) weights."""
        h = self.history[self.global_step - self.o.interval]
        h_rot, h_shape, h_cam = hmr_ref.forward(h['image'], self.with_buffers(params))
        h_s3d, _ = self.decode(h_rot, h_shape)
        h_s2d = self.project(h_cam, h_s3d)
        pred_motion = s2d_gt_part - h_s2d[:, 25:]
        gt_motion = kp_part[:, :, :-1] - h['s2d'][:, 25:, :-1]
        conf = ((h['s2d'][:, 25:, -1] + kp_part[:, :, -1]) == 2).to(self.dtype).unsqueeze(-1)
        loss = (((pred_motion - gt_motion) ** 2) * conf).mean()
        self.fit_losses['ul/motion_loss'] = loss
        return loss

    def labelled_loss(self, params, batch, tag):
        """:346-376."""
        conf = batch['keypoints'][:, 25:, -1].unsqueeze(-1)
        rot, shape, cam = hmr_ref.forward(batch['img'], self.with_buffers(params))
        s3d, _ = self.decode(rot, shape)
        gt_rot = G.batch_rodrigues(batch['pose'].view(-1, 3)).view(-1, 24, 3, 3)
        lpo, lsh = F.mse_loss(rot, gt_rot), F.mse_loss(shape, batch['betas'])
        s2d = self.project(cam, s3d)
        l2d = (((s2d[:, 25:] - batch['keypoints'][:, 25:, :-1]) ** 2) * conf).mean()
        l3d = self.s3d_loss(s3d[:, 25:], batch['pose_3d'][:, :, :-1], conf)
        loss = l2d * 5 + l3d * 5 + lsh * 0.001 + lpo * 1
        self.fit_losses.update({f'{tag}/labled_s2dloss': l2d, f'{tag}/labled_s3dloss': l3d,
                                f'{tag}/labled_shape_loss': lsh, f'{tag}/labled_pose_loss': lpo,
                                f'{tag}/labled_loss': loss})
        return loss

    def retrieval(self, feature):
        """:82-96 with the intended semantics for sample_num > 1 (tensor keys concatenated)."""
        dists = 1 - F.cosine_similarity(feature, self.centers)
        cluster = int(torch.argsort(dists)[0].item())
        picks = random.sample(self.index[cluster], self.o.sample_num)
        self.last_retrieval = (cluster, picks)
        idx = torch.as_tensor(picks, dtype=torch.long)
        return {k: v[idx] for k, v in self.bank.items()}

    # ------------------------------------------------------------------ levels
    def _level(self, image, kp, params, lower):
        o = self.o
        tag = 'll' if lower else 'ul'
        rot, shape, cam, feats = hmr_ref.forward(image, self.with_buffers(params), need_feature=True)
        s3d, _ = self.decode(rot, shape)
        s2d = self.project(cam, s3d)
        use_frame = o.use_frame_losses_lower if lower else o.use_frame_losses_upper
        use_temporal = o.use_temporal_losses_lower if lower else o.use_temporal_losses_upper
        loss = None
        if use_frame:
            loss, s2dloss = self.frame_losses(s2d, kp, shape, rot, tag)
            if lower:
                self.kp2dlosses_lower.append(s2dloss.item())
            else:
                self.kp2dlosses_upper[self.global_step] = s2dloss.item()
        if use_temporal:
            if o.use_meanteacher:
                tl = self.teacher_loss(image, rot, shape, s2d, s3d) * o.teacherloss_weight
                loss = tl if loss is None else loss + tl
            if o.use_motion and (self.global_step - o.interval) > 0:
                loss = loss + self.motion_loss(params, s2d[:, 25:], kp[:, 25:]) * o.motionloss_weight
        batch = None
        if o.retrieval:
            batch = self.retrieval(feats[5])
        if (o.lower_level_mixtrain if lower else o.upper_level_mixtrain):
            loss = loss + self.labelled_loss(params, batch, tag) * o.labelloss_weight
        return loss, feats

    def lower_level(self, image, kp, params):
        return self._level(image, kp, params, True)

    def upper_level(self, image, kp, params):
        return self._level(image, kp, params, False)

    def update_teacher(self):
        a = self.o.alpha                                                             # :193-201
        with torch.no_grad():
            for k, p in self.theta.items():
                self.teacher[k].mul_(a).add_(p.data, alpha=1 - a)

    def feature_diff(self, fa, fb):
        if self.cos_hook is not None:
            terms = torch.stack([torch.stack([(a.flatten().double() * b.flatten().double()).sum(), (a.flatten().double() ** 2).sum(),
                                              (b.flatten().double() ** 2).sum()]) for a, b in zip(fa, fb)])
            terms = self.cos_hook(terms)                                             # summed over the ranks
            sims = list((terms[:, 0] / (terms[:, 1].sqrt().clamp_min(1e-12) * terms[:, 2].sqrt().clamp_min(1e-12))).float())
        else:
            sims = [F.cosine_similarity(a.flatten(), b.flatten(), dim=0, eps=1e-12) for a, b in zip(fa, fb)]
        self.fit_losses['feat_sim/cos_sim'] = sum(sims) / (len(sims) - 1)          # :218 divides by last index
        return [s.item() for s in sims]

    # ------------------------------------------------------------------ driver
    def outer_step(self, loss):
        self.optimizer.zero_grad()
        loss.backward()
        if self.grad_hook is not None:
            self.grad_hook(self)
        self.optimizer.step()
        if self.o.use_meanteacher:
            self.update_teacher()

    def adaptation(self, batch, with_inference=True):
        """reference dynaboa_benchmark.py:126-201.  Returns a record of everything a parity test
        compares.  ``with_inference=False`` is the S-adapt scope."""
        o = self.o
        rec = {'lower_losses': [], 'cos': [], 'metrics': []}
        image, kp = batch['image'].to(self.dtype), batch['smpl_j2d'].to(self.dtype)
        self.history[self.global_step] = {'image': image.clone(), 's2d': kp.clone()}   # save_hist
        if not o.use_boa:
            loss, _ = self.lower_level(image, kp, self.theta)
            self.optimizer.zero_grad(); loss.backward(); self.optimizer.step()
            rec['upper_loss'] = loss.item()
            rec['metrics'].append(self.inference(batch, self.theta))
            return rec
        with torch.no_grad():
            init_feats = hmr_ref.forward(image, self.with_buffers(self.theta), need_feature=True)[3]
        fast = l2l_ref.clone_params(self.theta)
        for _ in range(o.inner_step):
            lloss, _ = self.lower_level(image, kp, fast)
            rec['lower_losses'].append(lloss.item())
            fast = l2l_ref.adapt_params(fast, lloss, o.fastlr, first_order=True)
            if with_inference:
                rec['metrics'].append(self.inference(batch, fast))
        uloss, _ = self.upper_level(image, kp, fast)
        rec['upper_loss'] = uloss.item()
        self.optimizer.zero_grad()
        uloss.backward()
        if self.grad_hook is not None:
            self.grad_hook(self)
        rec['grad_sample'] = {k: self.theta[k].grad.detach().clone() for k in
                              ('conv1.weight', 'layer2.0.bn1.weight', 'layer4.2.conv3.weight', 'fc1.bias',
                               'decpose.weight')}
        self.optimizer.step()
        if o.use_meanteacher:
            self.update_teacher()
        if with_inference:
            rec['metrics'].append(self.inference(batch, self.theta))
        rec['dynamic_steps'] = 0
        if o.dynamic_boa:
            with torch.no_grad():
                ad_feats = hmr_ref.forward(image, self.with_buffers(self.theta), need_feature=True)[3]
            cos = self.feature_diff(init_feats, ad_feats)
            rec['cos'].append(cos)
            steps = 0
            while 1 - cos[12] > o.cos_sim_threshold:
                steps += 1
                if steps > o.optim_steps:
                    break
                uloss, ad_feats = self.upper_level(image, kp, self.theta)
                self.outer_step(uloss)
                with torch.no_grad():
                    init_feats = [f.detach() for f in ad_feats]
                    ad_feats = hmr_ref.forward(image, self.with_buffers(self.theta), need_feature=True)[3]
                    cos = self.feature_diff(init_feats, ad_feats)
                    rec['cos'].append(cos)
                if with_inference:
                    rec['metrics'].append(self.inference(batch, self.theta))
            rec['dynamic_steps'] = steps
        return rec

    def predict(self, image, params=None):
        """Final outputs named by BASELINE.json: rotmat, betas, cam, joints49, vertices."""
        params = self.theta if params is None else params
        with torch.no_grad():
            rot, shape, cam = hmr_ref.forward(image.to(self.dtype), self.with_buffers(params))
            s3d, verts = self.decode(rot, shape)
        return dict(rotmat=rot, betas=shape, cam=cam, joints=s3d, vertices=verts)

    def inference(self, batch, params):
        """reference dynaboa_benchmark.py:204-262 -> (mpjpe, pampjpe, pve) in mm."""
        image = batch['image'].to(self.dtype)
        gt_pose, gt_betas, gender = batch['pose'].to(self.dtype), batch['betas'].to(self.dtype), batch['gender']
        with torch.no_grad():
            rot, shape, cam = hmr_ref.forward(image, self.with_buffers(params))
            _, verts = self.decode(rot, shape)
            _, gt_v = self.decode(gt_pose, gt_betas, 'male', pose2rot=True)
            _, gt_vf = self.decode(gt_pose, gt_betas, 'female', pose2rot=True)
            gt_v[gender == 1] = gt_vf[gender == 1]
            gt_k = torch.matmul(self.J_h36m, gt_v)
            gt_k = gt_k[:, self.h36m_to_j14] - gt_k[:, [0]]
            pr_k = torch.matmul(self.J_h36m, verts)
            pr_k = pr_k[:, self.h36m_to_j14] - pr_k[:, [0]]
            mpjpe = torch.sqrt(((pr_k - gt_k) ** 2).sum(-1)).mean(-1).numpy()
            S1, S2 = pr_k.numpy(), gt_k.numpy()
            S1_hat = np.stack([similarity_transform(a, b) for a, b in zip(S1, S2)])
            pampjpe = np.sqrt(((S1_hat - S2) ** 2).sum(-1)).mean(-1)
            _, gt_vn = self.decode(gt_pose, gt_betas, 'neutral', pose2rot=True)
            pve = np.sqrt(((gt_vn.numpy() - verts.numpy()) ** 2).sum(2)).mean()
        return mpjpe * 1000, pampjpe * 1000, pve * 1000
