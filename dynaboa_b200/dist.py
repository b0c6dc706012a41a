"""Data-parallel adaptation over streaming frames: one process per GPU, every rank holds a
full replica of theta / Adam state / teacher and adapts on its own frame (or its own video stream); the only
exchange step is ONE all-reduce (mean) of the flat fp32 outer gradient (107.9 MB) per outer update, plus a
3-float all-reduce so that every rank takes the same ``dynamic_boa`` branch.

The reference has no distributed code; this is the B200-native addition.  Semantics: per-rank inner SGD step,
rank-averaged outer gradient -- NOT identical to the reference at ``--batch_size R`` (which takes one inner step
on the batch-mean loss); ``mode='replicas'`` (no collective, R independent videos) is the exact-semantics mode.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of ``n_items`` for ``rank`` (first ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def make_grad_sync(world, group=None):
    """Hook for ``FusedAdam.pre_step_hook``: all-reduce(sum) of the flat gradient arena, then 1/world."""
    if world <= 1:
        return None

    def sync(flat_grad):
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad.mul_(1.0 / world)
    return sync


def allreduce_cosine_terms(dot, na2, nb2, group=None):
    """Sum (a.b, |a|^2, |b|^2) over ranks so that the dynamic-loop decision is identical everywhere and equals
    the reference's batch-level ``cal_feature_diff`` (it flattens across the batch, base_adaptor.py:215)."""
    t = torch.stack([dot, na2, nb2]) if torch.is_tensor(dot) else torch.tensor([dot, na2, nb2], dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t[0] / (t[1].sqrt().clamp_min(1e-12) * t[2].sqrt().clamp_min(1e-12))


def max_over_ranks(value, device=None):
    """MAX all-reduce of a python float (timings are reported as the max over ranks)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
