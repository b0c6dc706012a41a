"""Where one adapted frame (workload C2) spends its time on the CALLER's stream: CUDA events at the phase boundaries of
fused_adapt (probe forward, lower level, inner SGD, upper level, Adam), averaged over a few frames.  Side streams
(teacher forward, weight gradients, output forward) show up only as the time the caller's stream waits for them.

    python scripts/phase_times.py
"""
import os
import sys
import tempfile
from collections import OrderedDict

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import WORKLOADS, default_options  # noqa: E402
from dynaboa_b200 import config, synthetic  # noqa: E402
from dynaboa_b200.adaptor import Adaptor  # noqa: E402

work = tempfile.mkdtemp(prefix='dboa_phase_')
synthetic.write_asset_dir(os.path.join(work, 'data'))
config.set_data_root(os.path.join(work, 'data'))
N = 24
opts = default_options(expdir=work, expname='phase', model_file=config.BASE_MODEL, synthetic_frames=N, **WORKLOADS['c2'])
ad = Adaptor(opts)
ad.fused_eval = 'none'
stream = synthetic.SyntheticStream(length=N, batch_size=1)
frames = [{k: v.cuda() if torch.is_tensor(v) else v for k, v in stream[t].items()} for t in range(N)]
acc, cnt = OrderedDict(), 0
for t in range(N):
    ad.global_step, ad.fit_losses = t, {}
    ad.phase_events = [] if t >= 10 else None
    ad.adapt(frames[t])
    pred, ev = ad.predict_async(frames[t]['image'])
    if ad.phase_events is not None:
        torch.cuda.synchronize()
        evs = ad.phase_events
        for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
            acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
        cnt += 1
tot = sum(acc.values())
print(f'phases on the adaptation stream, mean of {cnt} frames (each frame synchronised: no overlap with the next frame); total {tot / cnt:.3f} ms')
for k, v in acc.items():
    print(f'  {k:45s} {v / cnt * 1e3:8.1f} us  {v / tot:6.1%}')
