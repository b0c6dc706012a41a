"""Oracle restatement of learn2learn 0.1.5 MAML (test infrastructure only).

``learn2learn==0.1.5`` is a third-party dependency of the reference (requirements.txt:10), not
vendored and not installable offline.  PARITY UNPINNED: restated from the published upstream
algorithm (learn2learn/algorithms/maml.py::{MAML.clone, MAML.adapt, maml_update},
learn2learn/utils::{clone_module, update_module}), anchored on the reference
call sites base_adaptor.py:119 (``MAML(model, lr=fastlr, first_order=True)``) and
dynaboa_benchmark.py:136,140 (``clone`` / ``adapt``).

Two forms:
* functional (``clone_params`` / ``adapt_params``) on ``name -> tensor`` dicts, used by
  ``adaptor_ref``;
* ``MAML`` nn.Module with the upstream interface, used by ``make_golden.py`` to stand in for the
  missing package when the reference's own ``BaseAdaptor`` code is executed on CPU.
"""
import torch
from torch import nn


# ---------------------------------------------------------------- functional form
def clone_params(params):
    """clone_module: every parameter becomes ``p.clone()`` (d clone / d p = I, so gradients of the
    fast weights reach the original leaves)."""
    return {k: v.clone() for k, v in params.items()}


def adapt_params(fast, loss, lr, first_order=True):
    """MAML.adapt + maml_update: g = grad(loss, fast); fast <- fast + (-lr * g)."""
    names = list(fast.keys())
    grads = torch.autograd.grad(loss, [fast[k] for k in names], retain_graph=not first_order,
                                create_graph=not first_order, allow_unused=False)
    return {k: fast[k] + (-lr * g) for k, g in zip(names, grads)}


# ---------------------------------------------------------------- module form
def _clone_module(module, memo=None):
    if memo is None:
        memo = {}
    clone = module.__new__(type(module))
    clone.__dict__ = module.__dict__.copy()
    clone._parameters = clone._parameters.copy()
    clone._buffers = clone._buffers.copy()
    clone._modules = clone._modules.copy()
    for key, p in module._parameters.items():
        if p is not None:
            ptr = p.data_ptr()
            if ptr not in memo:
                memo[ptr] = p.clone()
            clone._parameters[key] = memo[ptr]
    for key, b in module._buffers.items():
        if b is not None and b.requires_grad:
            clone._buffers[key] = b.clone()
    for key, child in module._modules.items():
        clone._modules[key] = _clone_module(child, memo)
    return clone


def _update_module(module, memo=None):
    if memo is None:
        memo = {}
    for key, p in module._parameters.items():
        if p is not None and getattr(p, 'update', None) is not None:
            if p in memo:
                module._parameters[key] = memo[p]
            else:
                new = p + p.update
                memo[p] = new
                module._parameters[key] = new
    for child in module._modules.values():
        _update_module(child, memo)
    return module


class MAML(nn.Module):
    def __init__(self, model, lr, first_order=False, allow_unused=None, allow_nograd=False):
        super().__init__()
        self.module = model
        self.lr = lr
        self.first_order = first_order
        self.allow_unused = allow_nograd if allow_unused is None else allow_unused
        self.allow_nograd = allow_nograd

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def adapt(self, loss, first_order=None):
        first_order = self.first_order if first_order is None else first_order
        params = list(self.module.parameters())
        grads = torch.autograd.grad(loss, params, retain_graph=not first_order,
                                    create_graph=not first_order, allow_unused=self.allow_unused)
        for p, g in zip(params, grads):
            p.update = None if g is None else -self.lr * g
        self.module = _update_module(self.module)

    def clone(self, first_order=None):
        first_order = self.first_order if first_order is None else first_order
        return MAML(_clone_module(self.module), lr=self.lr, first_order=first_order,
                    allow_unused=self.allow_unused, allow_nograd=self.allow_nograd)
