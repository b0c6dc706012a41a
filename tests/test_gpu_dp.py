"""Data-parallel adaptation (dynaboa_b200/dist.py: bucketed all-reduce of the outer gradient under the upper-level backward,
1/world folded into Adam, rank-summed cosine terms for the dynamic loop) against the CPU emulation of R ranks
(oracle/dp_ref.py, fixture tests/golden/adapt_dp2.npz): two processes, one frame stream each.  With two GPUs the ranks use NCCL
on their own device; on a one-GPU box they share the device and exchange through gloo (same host logic, same kernels)."""
import ast
import os
import socket
import tempfile
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, 'tests', 'golden', 'adapt_dp2.npz')


def _rank_main(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    import torch.distributed as dist
    from dynaboa_b200 import config, dist as ddist, synthetic
    from dynaboa_b200.adaptor import Adaptor
    from oracle.make_golden import sample_indices
    multi = torch.cuda.device_count() >= world
    torch.cuda.set_device(rank if multi else 0)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('nccl' if multi else 'gloo', rank=rank, world_size=world)
    gd = dict(np.load(GOLDEN, allow_pickle=False))
    work = tempfile.mkdtemp(prefix=f'dboa_dp_r{rank}_')
    synthetic.write_asset_dir(os.path.join(work, 'data'))
    config.set_data_root(os.path.join(work, 'data'))
    o = dict(ast.literal_eval(str(gd['options'])))
    n_frames = gd['dyn_steps'].shape[1]
    o.update(expdir=work, expname='dp', tensorboard=0, synthetic_frames=n_frames, cache_results=0, model_file=config.BASE_MODEL, rank=rank,
             dataset='3dpw', save_res=0, seq_seed=22)
    ad = Adaptor(SimpleNamespace(**o))
    ad.fused_eval = 'none'
    ddist.attach(ad, world)
    stream = synthetic.SyntheticStream(length=n_frames, batch_size=1, rank=rank)
    for t in range(n_frames):
        batch = {k: v.cuda() if torch.is_tensor(v) else v for k, v in stream[t].items()}
        ad.global_step, ad.fit_losses = t, {}
        ad.adapt(batch)
        assert ad.optim_step_record[-1] == int(gd['dyn_steps'][rank][t]), (rank, t, ad.optim_step_record)
        assert abs(float(ad.last_upper_loss) - gd['upper_loss'][rank][t]) <= 1e-3 * abs(gd['upper_loss'][rank][t]), (rank, t)
    pred = ad.predict(batch['image'])
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.double().cpu() - torch.as_tensor(b).double()).abs().max() / torch.as_tensor(b).double().abs().max())
    for key in ('rotmat', 'betas', 'cam', 'joints'):
        assert rel(pred[key], gd[f'{key}_r{rank}']) < 1e-3, (rank, key)
    n_outer = int((1 + np.minimum(gd['dyn_steps'][rank], o['optim_steps'])).sum())
    params = dict(ad.model.module.named_parameters())
    for i, name in enumerate(str(s) for s in gd['param_names']):
        p = params[name]
        th = p.detach().contiguous().flatten()[sample_indices(name, p.numel())].double().cpu().numpy()
        assert np.abs(th - gd['theta_samples'][i]).max() <= 4 * o['lr'] * n_outer, (rank, name)
    # replicas stay identical: compare a digest of theta across the ranks
    digest = torch.stack([ad.model.module.arena.double().sum(), ad.model.module.arena.double().abs().sum()]).cuda()
    both = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(both, digest)
    assert torch.equal(both[0], both[1]), 'replicas diverged'
    open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    dist.destroy_process_group()


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason='tests/golden/adapt_dp2.npz missing')
def test_two_ranks_follow_the_rank_emulation(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / 'ok0') and os.path.exists(tmp_path / 'ok1')
