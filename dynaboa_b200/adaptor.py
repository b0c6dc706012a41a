"""Benchmark driver on top of ``BaseAdaptor`` (the role of reference dynaboa_benchmark.py ``Adaptor``
:69-262: ``excute`` frame loop, ``adaptation`` bilevel + dynamic loop, ``inference`` metrics).

Two entry points per frame:

* ``adaptation(batch)``    -- the reference's control flow through torch autograd (``learner.adapt``,
  ``loss.backward()``, ``optimizer.step()``), every op a libdynaboa_b200 kernel.  This is the path the
  unchanged reference driver also exercises.
* ``adapt(batch)``         -- the same arithmetic without autograd: direct C-ABI calls into one flat gradient
  arena, Adam and the teacher EMA fused in one sweep, no host syncs except the ``dynamic_boa`` decision
  (see fused.py).  This is what ``bench.py`` times.
"""
import os.path as osp

import numpy as np
import torch

from . import _lib, constants
from .base_adaptor import BaseAdaptor


class Adaptor(BaseAdaptor):
    def __init__(self, options):
        super().__init__(options)
        self.model.eval()        # the reference driver does this before every frame (dynaboa_benchmark.py:89)
        self.reset_records()

    def reset_records(self):
        n = len(self.dataloader)
        self.feat_sims, self.optim_step_record = {}, []
        self.mpjpe_statistics, self.pampjpe_statistics = [[] for _ in range(n)], [[] for _ in range(n)]
        self.mpjpe_all_lower = [[] for _ in range(self.options.inner_step)]
        self.pampjpe_all_lower = [[] for _ in range(self.options.inner_step)]
        self.history, self.kp2dlosses_lower, self.kp2dlosses_upper = {}, [], {}

    # ------------------------------------------------------------------ frame loop (reference :71-123)
    def excute(self, max_frames=None, fused=False):
        self.reset_records()
        mpjpe_all, pampjpe_all, pve_all = [], [], []
        for step, batch in enumerate(self.dataloader):
            if max_frames is not None and step >= max_frames:
                break
            self.global_step, self.fit_losses = step, {}
            batch = {k: v.to(self.device) if isinstance(v, torch.Tensor) else v for k, v in batch.items()}
            self.model.eval()
            mpjpe, pampjpe, pve = (self.adapt(batch) if fused else self.adaptation(batch))
            self.fit_losses.update({'metrics/mpjpe': mpjpe, 'metrics/pampjpe': pampjpe, 'metrics/pve': pve})
            self.write_summaries(self.fit_losses)
            mpjpe_all.append(mpjpe); pampjpe_all.append(pampjpe); pve_all.append(pve)
        summary = dict(mpjpe=float(np.mean(mpjpe_all)), pampjpe=float(np.mean(pampjpe_all)), pve=float(np.mean(pve_all)))
        self.save_records(mpjpe_all, pampjpe_all, pve_all)
        return summary

    def save_records(self, mpjpe_all, pampjpe_all, pve_all):
        """The eight result files of the reference driver, same names, same joblib pickles, same keys
        (dynaboa_benchmark.py:111-123)."""
        import joblib
        p = lambda name: osp.join(self.exppath, name)
        joblib.dump({'kp2dloss': [float(x) for x in self.kp2dlosses_lower]}, p('lowerlevel_kp2dloss.pt'))
        joblib.dump({'kp2dloss': {k: float(v) for k, v in self.kp2dlosses_upper.items()}}, p('upperlevel_kp2dloss.pt'))
        joblib.dump({'mpjpe': mpjpe_all, 'pampjpe': pampjpe_all, 'pve': pve_all}, p('res.pt'))
        joblib.dump({'mpjpe': self.mpjpe_all_lower, 'pampjpe': self.pampjpe_all_lower}, p('lower_res.pt'))
        joblib.dump({'mpjpe': self.mpjpe_statistics, 'pampjpe': self.pampjpe_statistics}, p('steps_statistic_res.pt'))
        joblib.dump({'feat': self.feat_sims}, p('feat_sims.pt'))
        joblib.dump({'step': self.optim_step_record}, p('optim_step_record.pt'))
        with open(p('res.txt'), 'w') as f:
            f.write(f'Step:{self.global_step}: MPJPE:{np.mean(mpjpe_all)}, PAMPJPE:{np.mean(pampjpe_all)}, PVE:{np.mean(pve_all)}\n')
            for i in range(self.options.inner_step):
                lo_m = np.mean(self.mpjpe_all_lower[i]) if len(self.mpjpe_all_lower[i]) else float('nan')
                lo_p = np.mean(self.pampjpe_all_lower[i]) if len(self.pampjpe_all_lower[i]) else float('nan')
                f.write(f'Lower-level  Step:{i} MPJPE:{lo_m}, PAMPJPE:{lo_p}\n')

    def _outer_step(self, loss):
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        if self.options.use_meanteacher:
            self.update_teacher(self.teacher, self.model)

    # ------------------------------------------------------------------ bilevel step through autograd (reference :126-201)
    def adaptation(self, batch):
        o = self.options
        image, kp = batch['image'], batch['smpl_j2d']
        self.save_hist(image, kp)
        if not o.use_boa:
            loss, _ = self.lower_level_adaptation(image, kp, None, self.model)
            self.optimizer.zero_grad(); loss.backward(); self.optimizer.step()
            return self.inference(batch, self.model)

        with torch.no_grad():
            init_features = self.model(image, need_feature=True)[3]
        learner = self.model.clone()
        for i in range(o.inner_step):
            lower_loss, _ = self.lower_level_adaptation(image, kp, None, learner)
            learner.adapt(lower_loss)
            mpjpe, pampjpe, _ = self.inference(batch, learner)
            self.fit_losses[f'metrics/lower_{i}_mpjpe'], self.fit_losses[f'metrics/lower_{i}_pampjpe'] = mpjpe, pampjpe
            self.mpjpe_all_lower[i].append(mpjpe); self.pampjpe_all_lower[i].append(pampjpe)
        upper_loss, _ = self.upper_level_adaptation(image, kp, None, learner)
        self.last_upper_loss = upper_loss.detach()
        self._outer_step(upper_loss)
        mpjpe, pampjpe, pve = self.inference(batch, self.model)
        if self.global_step < len(self.mpjpe_statistics):
            self.mpjpe_statistics[self.global_step], self.pampjpe_statistics[self.global_step] = [mpjpe], [pampjpe]

        if o.dynamic_boa:
            with torch.no_grad():
                adapted = self.model(image, need_feature=True)[3]
                sims = self.cal_feature_diff(init_features, adapted)
            self.feat_sims[self.global_step] = [sims]
            self.optimized_step = 0
            while 1 - sims[12]['cos'] > o.cos_sim_threshold:
                self.optimized_step += 1
                if self.optimized_step > o.optim_steps:
                    break
                upper_loss, adapted = self.upper_level_adaptation(image, kp, None, self.model)
                self._outer_step(upper_loss)
                with torch.no_grad():
                    init_features = adapted
                    adapted = self.model(image, need_feature=True)[3]
                    sims = self.cal_feature_diff(init_features, adapted)
                self.feat_sims[self.global_step].append(sims)
                mpjpe, pampjpe, pve = self.inference(batch, self.model)
            self.optim_step_record.append(self.optimized_step)
        return mpjpe, pampjpe, pve

    # ------------------------------------------------------------------ evaluation (reference :204-262)
    def predict(self, image, model=None):
        """rotmat, betas, cam, joints49, vertices for ``image`` (no adaptation)."""
        model = self.model if model is None else model
        with torch.no_grad():
            rot, shape, cam = model(image)
            out = self.decode_smpl_params(rot, shape)
        return dict(rotmat=rot, betas=shape, cam=cam, joints=out['s3d'], vertices=out['vts'])

    def predict_async(self, image):
        """``predict`` issued on a side stream.  The output forward of frame t and the adaptation of frame t+1 read the
        same weights, so the caller can go on enqueueing the next frame on its own stream: the two chains of small kernels
        overlap on the device, and the next optimiser step waits for this read before it overwrites the weights.
        Returns (outputs, event); consumers on another stream (or the host) wait for the event first."""
        cur = torch.cuda.current_stream()
        if getattr(self, 'output_stream', None) is None:
            self.output_stream = torch.cuda.Stream(device=image.device)
        side = self.output_stream
        side.wait_stream(cur)                              # the weights (and the frame) are ready
        with torch.cuda.stream(side):
            out = self.predict(image)
            ev = torch.cuda.Event()
            ev.record(side)
        image.record_stream(side)
        for v in out.values():                             # allocated on the side stream, read by the caller's stream after the event
            v.record_stream(cur)
        self.optimizer.wait_before_write.append(ev)
        return out, ev

    def inference(self, batch, model, need_feature=False):
        image, gt_pose, gt_betas, gender = batch['image'], batch['pose'], batch['betas'], batch['gender']
        model.eval()
        with torch.no_grad():
            out = model(image, need_feature)
            pred_rotmat, pred_shape, pred_cam = out[0], out[1], out[2]
            pred_vertices = self.decode_smpl_params(pred_rotmat, pred_shape)['vts']
            gt_vertices = self.smpl_male(global_orient=gt_pose[:, :3], body_pose=gt_pose[:, 3:], betas=gt_betas).vertices
            gt_vertices_f = self.smpl_female(global_orient=gt_pose[:, :3], body_pose=gt_pose[:, 3:], betas=gt_betas).vertices
            gt_vertices = torch.where((gender == 1).view(-1, 1, 1), gt_vertices_f, gt_vertices)
            gt_neutral = self.smpl_neutral(betas=gt_betas, body_pose=gt_pose[:, 3:], global_orient=gt_pose[:, :3], pose2rot=True).vertices
            # H36M-regressor joints, MPJPE, Procrustes PA-MPJPE and PVE in two launches; ONE 3-float read-back per sample
            # (the reference copies joints and both meshes to the host and runs a numpy SVD per sample: :217-240)
            metrics = self.eval_metrics(pred_vertices, gt_vertices, gt_neutral).cpu().numpy()
            mpjpe, pampjpe, pve = metrics[:, 0], metrics[:, 1], float(metrics[:, 2].mean())
        if getattr(self.options, 'cache_results', 1):        # the reference always dumps Pred_<step>.pt (:250-254); 0 turns the disk write off
            cam_t = torch.stack([pred_cam[:, 1], pred_cam[:, 2], 2 * 5000. / (constants.IMG_RES * pred_cam[:, 0] + 1e-9)], dim=-1)
            import joblib
            joblib.dump({'verts': pred_vertices.cpu().numpy(), 'cam': cam_t.cpu().numpy(), 'rotmat': pred_rotmat.cpu().numpy(),
                         'beta': pred_shape.cpu().numpy()}, osp.join(self.exppath, 'result', f'Pred_{self.global_step}.pt'))
        if need_feature:
            return mpjpe * 1000, pampjpe * 1000, pve * 1000, out[3]
        return mpjpe * 1000, pampjpe * 1000, pve * 1000

    def eval_metrics(self, pred_vertices, gt_vertices, gt_vertices_neutral):
        """(B,3) device tensor: MPJPE, PA-MPJPE, PVE (metres) of each sample -- ``dboa_eval_metrics``."""
        B, dev = pred_vertices.shape[0], pred_vertices.device
        if getattr(self, '_eval_consts', None) is None or self._eval_consts[0].device != dev:
            self._eval_consts = (self.J_regressor.to(dev).contiguous().float(),
                                 torch.as_tensor(np.asarray(self.joint_mapper_h36m), dtype=torch.int32, device=dev))
        J, jmap = self._eval_consts
        _lib.require_cuda(pred_vertices, gt_vertices, gt_vertices_neutral)
        scratch = torch.empty(_lib.load().dboa_eval_scratch_floats(B, J.shape[0]), dtype=torch.float32, device=dev)
        out = torch.empty(B, 3, dtype=torch.float32, device=dev)
        keep = [t.contiguous().float() for t in (pred_vertices, gt_vertices, gt_vertices_neutral)]
        _lib.call('dboa_eval_metrics', _lib.ptr(keep[0]), _lib.ptr(keep[1]), _lib.ptr(keep[2]), _lib.ptr(J), int(J.shape[0]),
                  int(pred_vertices.shape[1]), _lib.ptr(jmap), int(jmap.numel()), _lib.ptr(scratch), _lib.ptr(out), B, _lib.stream())
        return out

    def adapt(self, batch):
        from .fused import fused_adapt
        return fused_adapt(self, batch)


class InternetAdaptor(Adaptor):
    """Client of the same path without ground truth (reference dynaboa_internet.py ``Adaptor`` :68-168): the adaptation is
    ``Adaptor.adaptation`` minus the per-step evaluation, ``inference`` only decodes the mesh and caches ``Pred_<step>.pt``."""

    def reset_records(self):
        self.feat_sims, self.optim_step_record = {}, []
        self.mpjpe_statistics, self.pampjpe_statistics = [], []
        self.mpjpe_all_lower = [[] for _ in range(self.options.inner_step)]
        self.pampjpe_all_lower = [[] for _ in range(self.options.inner_step)]
        self.history, self.kp2dlosses_lower, self.kp2dlosses_upper = {}, [], {}

    def excute(self, max_frames=None, fused=False):
        self.reset_records()
        outs = []
        for step, batch in enumerate(self.dataloader):
            if max_frames is not None and step >= max_frames:
                break
            self.global_step, self.fit_losses = step, {}
            batch = {k: v.to(self.device) if isinstance(v, torch.Tensor) else v for k, v in batch.items()}
            self.model.eval()
            if fused:
                self.fused_eval = 'none'
                self.adapt(batch)
            else:
                self.adaptation(batch)
            outs.append(self.inference(batch, self.model))
        return outs

    def inference(self, batch, model, need_feature=False):
        """reference dynaboa_internet.py:142-168: no metrics (there is no ground truth), cached predictions only."""
        model.eval()
        with torch.no_grad():
            out = model(batch['image'], need_feature)
            pred_rotmat, pred_shape, pred_cam = out[0], out[1], out[2]
            pred_vertices = self.decode_smpl_params(pred_rotmat, pred_shape)['vts']
        cam_t = torch.stack([pred_cam[:, 1], pred_cam[:, 2], 2 * 5000. / (constants.IMG_RES * pred_cam[:, 0] + 1e-9)], dim=-1)
        res = {'verts': pred_vertices, 'cam': cam_t, 'rotmat': pred_rotmat, 'beta': pred_shape}
        if getattr(self.options, 'cache_results', 1):
            import joblib
            joblib.dump({k: v.cpu().numpy() for k, v in res.items()}, osp.join(self.exppath, 'result', f'Pred_{self.global_step}.pt'))
        zero = np.zeros(pred_rotmat.shape[0], np.float32)
        return (zero, zero, 0.0, out[3]) if need_feature else (zero, zero, 0.0)
