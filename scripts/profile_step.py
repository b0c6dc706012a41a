"""Profiling harness: runs a few adapted frames of workload C2 and brackets ONE frame (or one HMR forward) with
cudaProfilerStart/Stop so that `ncu --profile-from-start off` captures exactly that region.

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python scripts/profile_step.py
"""
import argparse
import os
import sys
import tempfile

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

ap = argparse.ArgumentParser()
ap.add_argument('--region', default='frame', choices=['frame', 'forward', 'fwdbwd'])
ap.add_argument('--tc', type=int, default=0)
ap.add_argument('--batch', type=int, default=1)
args = ap.parse_args()

from bench import WORKLOADS, default_options  # noqa: E402
from dynaboa_b200 import _lib, config, hmr as hmr_mod, synthetic  # noqa: E402
from dynaboa_b200.adaptor import Adaptor  # noqa: E402

lib = _lib.load()
lib.dboa_set_tensor_core_conv(args.tc)
work = tempfile.mkdtemp(prefix='dboa_prof_')
synthetic.write_asset_dir(os.path.join(work, 'data'))
config.set_data_root(os.path.join(work, 'data'))
opts = default_options(expdir=work, expname='prof', model_file=config.BASE_MODEL, synthetic_frames=10, **WORKLOADS['c2'])
ad = Adaptor(opts)
ad.fused_eval = 'none'
stream = synthetic.SyntheticStream(length=10, batch_size=1)
frames = [{k: v.cuda() if torch.is_tensor(v) else v for k, v in stream[t].items()} for t in range(10)]
model = ad.model.module
if args.region == 'frame':
    for t in range(8):
        ad.global_step, ad.fit_losses = t, {}
        ad.adapt(frames[t])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ad.global_step = 8
    ad.adapt(frames[8])
    ad.predict(frames[8]['image'])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    x = frames[0]['image'].repeat(args.batch, 1, 1, 1)
    tape = torch.empty(hmr_mod.tape_floats(args.batch), device='cuda')
    G = torch.zeros_like(model.arena)
    for _ in range(2):
        rot, shape, cam, _, _ = hmr_mod.raw_forward(model.arena, model._buffers, x, None, tape)
        hmr_mod.raw_backward(model.arena, tape, args.batch, False, torch.ones_like(rot), torch.ones_like(shape), torch.ones_like(cam), G)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    rot, shape, cam, _, _ = hmr_mod.raw_forward(model.arena, model._buffers, x, None, tape)
    if args.region == 'fwdbwd':
        hmr_mod.raw_backward(model.arena, tape, args.batch, False, torch.ones_like(rot), torch.ones_like(shape), torch.ones_like(cam), G)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('profiled region done')
