"""Third client of the adaptation path: frames + OpenPose detections instead of a dataset (the role of reference
dynaboa_webcam.py ``Adaptor`` :36-337; the capture / OpenPose / viewer loop :373-447 stays with the caller).

What differs from the benchmark client, and where it lands here:

* the 2D re-projection and motion terms compare the 25 OpenPose joints ``pred_s2d[:, :25]`` with the detections
  (reference :236,246,262 and :161-181) instead of the 24 ground-truth joints 25..48: ``kp_range = (0, 25)`` selects the
  joints inside the loss kernels (``dboa_loss_args.kp_first / kp_count``, ``dboa_loss_motion_joints``); the detections
  travel in rows 0..24 of the usual (B, 49, 3) keypoint block;
* the lower level has the frame losses only, the upper level adds motion and mean-teacher terms (:228-271);
* the teacher runs in eval mode (:69) -- no dropout masks; no exemplar retrieval;
* ``save_hist`` advances ``global_step`` itself (:105-108), so the motion term looks ``interval - 1`` frames back and
  starts one frame earlier than in the benchmark driver;
* the input side (``dataprocess`` :185-206: bounding box from the detections, confidence threshold 0.3, crop / resize /
  normalise) runs on the GPU through dynaboa_b200.dataprocess.

``online_adaptation(frame, detections)`` returns the reference's result dictionary (``vts``, ``cam``, ``bbox`` and, with
``test_basemodel``, ``vts_base`` / ``cam_base``).
"""
from types import SimpleNamespace

import numpy as np
import torch

from . import config, dataprocess
from .adaptor import Adaptor
from .hmr import hmr


def webcam_options(o):
    """The webcam script's flag namespace (reference dynaboa_webcam.py:339-371) completed with the fields ``BaseAdaptor``
    reads, set to what the webcam ``Adaptor`` hard-codes."""
    d = dict(vars(o))
    d.update(use_frame_losses_lower=1, use_frame_losses_upper=1, use_temporal_losses_lower=0, use_temporal_losses_upper=1,
             retrieval=0, lower_level_mixtrain=0, upper_level_mixtrain=0, sample_num=1, inner_step=1, batch_size=1,
             teacher_dropout=0, dataset='webcam')
    for k, v in dict(labelloss_weight=0.0, tensorboard=0, expdir='exps', expname='webcam', use_meanteacher=0, use_motion=0,
                     dynamic_boa=0, cos_sim_threshold=3.1e-4, optim_steps=7, interval=5, alpha=0.1, test_basemodel=0,
                     motionloss_weight=0.8, teacherloss_weight=0.1, seed=22).items():
        d.setdefault(k, v)
    return SimpleNamespace(**d)


class WebcamAdaptor(Adaptor):
    kp_range = (0, 25)

    def __init__(self, options):
        super().__init__(webcam_options(options))
        self.fused_eval = 'none'
        self.basemodel = None
        if self.options.test_basemodel:                     # frozen copy of the base model, shown side by side (:71-74)
            ck = torch.load(self.options.model_file, map_location='cpu', weights_only=False)
            self.basemodel = hmr(config.SMPL_MEAN_PARAMS).to(self.device)
            self.basemodel.load_state_dict({k.replace('module.', ''): v for k, v in ck['model'].items()}, strict=True)
            self.basemodel.eval()

    def set_dataloader(self):
        self.dataloader, self.imgdir = [], None

    def save_hist(self, image, s2d):
        super().save_hist(image, s2d)
        self.global_step += 1                               # reference :105-108

    def reload(self):
        """Key ``r`` of the demo (:184-195): back to the base checkpoint, fresh Adam state."""
        self.set_model_optim()
        self.model.eval()
        if self.options.use_meanteacher:
            self.set_teacher()
        self.history = {}

    def dataprocess(self, image, gtkp2d, scaleFactor=1.0):
        """reference :197-218.  ``image``: (H, W, 3) RGB uint8 / float32 frame (numpy or CUDA tensor); ``gtkp2d``: (25, 3) OpenPose
        detections (x, y, confidence) in frame pixels.  Returns the (1, 3, 224, 224) network input, the (1, 49, 3) keypoint
        block (rows 0..24 filled) and the (1, 3) bounding box."""
        kp = np.array(gtkp2d, dtype=np.float64, copy=True)
        x0, y0, x1, y1 = kp[:, 0].min(), kp[:, 1].min(), kp[:, 0].max(), kp[:, 1].max()
        center = [(x1 + x0) / 2, (y1 + y0) / 2]
        scale = scaleFactor * max(x1 - x0, y1 - y0) / 200
        bbox = np.stack([center[0], center[1], scale * 200])
        kp[:, 2] = kp[:, 2] > 0.3
        img = image if torch.is_tensor(image) else torch.from_numpy(np.ascontiguousarray(image))
        img = img.to(self.device, non_blocking=True)
        net_in = dataprocess.crop(img, center, scale)
        kp25 = dataprocess.j2d_processing(torch.from_numpy(kp).float().to(self.device), center, scale)
        kp49 = torch.zeros(1, 49, 3, dtype=torch.float32, device=self.device)
        kp49[0, :25] = kp25
        return net_in.unsqueeze(0), kp49, bbox[None, :]

    def online_adaptation(self, ori_image, gtkp2d):
        """reference :221-337: one adaptation step on the frame, then the adapted model's mesh."""
        image, kp, bbox = self.dataprocess(ori_image, np.asarray(gtkp2d)[0], scaleFactor=1.2)
        return self.adapt_processed(image, kp, bbox)

    def adapt_processed(self, image, kp, bbox=None):
        """The same on an already cropped frame and a (B, 49, 3) keypoint block (rows 0..24 = OpenPose joints in [-1, 1])."""
        self.fit_losses = {}
        self.model.eval()
        self.adapt({'image': image, 'smpl_j2d': kp})        # fused path: bilevel step (+ dynamic loop), no evaluation
        with torch.no_grad():
            rot, shape, cam = self.model(image)
            res = {'vts': self.decode_smpl_params(rot, shape)['vts'], 'cam': cam, 'bbox': bbox}
            if self.basemodel is not None:
                rot_b, shape_b, cam_b = self.basemodel(image)
                res.update(vts_base=self.decode_smpl_params(rot_b, shape_b)['vts'], cam_base=cam_b)
        return res
