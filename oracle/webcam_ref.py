"""Oracle restatement of the webcam client's adaptation step (test infrastructure only; never imported by the product).

CPU torch restatement of reference dynaboa_webcam.py ``Adaptor.online_adaptation`` :221-337 (without the input processing
:197-218, restated in oracle/dataprocess_ref.py) on top of ``OracleAdaptor``: the 2D terms compare the 25 OpenPose joints
``pred_s2d[:, :25]`` with the detections (:236,246,262), the motion term does the same on the history frame (:161-181), the lower
level has the frame losses only, the upper level adds motion and the mean-teacher terms (:228-271), the teacher is in eval mode
(:69), ``save_hist`` advances ``global_step`` (:105-108).  Pinned against the reference class itself by
oracle/make_golden.py ``golden_webcam`` (the class body is executed unmodified on CPU).
"""
import torch

from . import hmr_ref, l2l_ref
from .adaptor_ref import OracleAdaptor, default_options


def webcam_options(**over):
    """Flag defaults of reference dynaboa_webcam.py:339-371 plus what its ``Adaptor`` hard-codes."""
    o = default_options(retrieval=0, lower_level_mixtrain=0, upper_level_mixtrain=0, inner_step=1, teacher_dropout=0,
                        use_frame_losses_lower=1, use_frame_losses_upper=1, use_temporal_losses_lower=0, use_temporal_losses_upper=1)
    for k, v in over.items():
        setattr(o, k, v)
    return o


class OracleWebcam(OracleAdaptor):
    def frame_losses(self, s2d, kp, shape, rotmat, tag):
        conf = kp[:, :, -1].unsqueeze(-1)                                            # :234-241
        s2dloss = (((s2d[:, :25] - kp[:, :, :-1]) ** 2) * conf).mean()
        sp, pp = self.shape_prior(shape), self.pose_prior(rotmat)
        return s2dloss * self.o.s2dloss_weight + sp * self.o.shape_prior_weight + pp * self.o.pose_prior_weight, s2dloss

    def motion_loss(self, params, s2d25, kp):
        h = self.history[self.global_step - self.o.interval]                        # :161-181
        h_rot, h_shape, h_cam = hmr_ref.forward(h['image'], self.with_buffers(params))
        h_s2d = self.project(h_cam, self.decode(h_rot, h_shape)[0])
        pred_motion = s2d25 - h_s2d[:, :25]
        gt_motion = kp[:, :, :-1] - h['s2d'][:, :, :-1]
        conf = ((h['s2d'][:, :, -1] + kp[:, :, -1]) == 2).to(self.dtype).unsqueeze(-1)
        return (((pred_motion - gt_motion) ** 2) * conf).mean()

    def _level(self, image, kp, params, lower):
        o = self.o
        rot, shape, cam, feats = hmr_ref.forward(image, self.with_buffers(params), need_feature=True)
        s3d, _ = self.decode(rot, shape)
        s2d = self.project(cam, s3d)
        loss, _ = self.frame_losses(s2d, kp, shape, rot, 'll' if lower else 'ul')
        if not lower:
            if o.use_motion and (self.global_step - o.interval) > 0:                # :263-265
                loss = loss + self.motion_loss(params, s2d[:, :25], kp) * o.motionloss_weight
            if o.use_meanteacher:                                                    # :266-268
                loss = loss + self.teacher_loss(image, rot, shape, s2d, s3d) * o.teacherloss_weight
        return loss, feats

    def online_adaptation(self, image, kp):
        """``image``: (B, 3, 224, 224) network input; ``kp``: (B, 25, 3) OpenPose joints in [-1, 1] + confidence.  Returns the
        number of dynamic re-adaptation steps taken."""
        o = self.o
        image, kp = image.to(self.dtype), kp.to(self.dtype)
        self.history[self.global_step] = {'image': image.clone(), 's2d': kp.clone()}
        self.global_step += 1
        if not o.use_boa:
            loss, _ = self._level(image, kp, self.theta, True)
            self.optimizer.zero_grad(); loss.backward(); self.optimizer.step()
            return 0
        with torch.no_grad():
            init_feats = hmr_ref.forward(image, self.with_buffers(self.theta), need_feature=True)[3]
        fast = l2l_ref.clone_params(self.theta)
        lloss, _ = self._level(image, kp, fast, True)
        fast = l2l_ref.adapt_params(fast, lloss, o.fastlr, first_order=True)
        uloss, _ = self._level(image, kp, fast, False)
        self.outer_step(uloss)
        steps = 0
        if o.dynamic_boa:
            with torch.no_grad():
                ad_feats = hmr_ref.forward(image, self.with_buffers(self.theta), need_feature=True)[3]
            cos = self.feature_diff(init_feats, ad_feats)
            while 1 - cos[12] > o.cos_sim_threshold:
                steps += 1
                if steps > o.optim_steps:
                    break
                uloss, _ = self._level(image, kp, self.theta, False)
                self.outer_step(uloss)
                with torch.no_grad():
                    init_feats, ad_feats = ad_feats, hmr_ref.forward(image, self.with_buffers(self.theta), need_feature=True)[3]
                    cos = self.feature_diff(init_feats, ad_feats)
        return steps
