"""First-order MAML wrapper with the learn2learn 0.1.5 interface used by the reference
(``l2l.algorithms.MAML(model, lr=fastlr, first_order=True)`` base_adaptor.py:119; ``clone`` / ``adapt``
dynaboa_benchmark.py:136,140).

Fast weights are one flat arena: ``clone()`` is a single differentiable copy whose backward routes the
outer gradient to the 169 leaf parameters unchanged (d clone / d theta = I), and ``adapt(loss)`` is one
``torch.autograd.grad`` on that flat tensor plus one fused sweep ``theta' = theta' + (-lr * g)``.  With
``first_order=True`` the inner gradient is detached, so the adjoint of the inner SGD update is the identity --
exactly what the reference computes; the second-order term is not part of the reference path and is rejected.
"""
import torch
from torch import nn


class MAML(nn.Module):
    def __init__(self, model, lr, first_order=False, allow_unused=None, allow_nograd=False):
        super().__init__()
        if not first_order:
            raise NotImplementedError('only first_order=True (the reference configuration) is implemented')
        self.module = model
        self.lr = lr
        self.first_order = first_order
        self.allow_unused = allow_nograd if allow_unused is None else allow_unused
        self.allow_nograd = allow_nograd

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    # A learner returned by clone() exposes its FAST weights (non-leaf views of the flat fast arena), as learn2learn's cloned
    # module does; nn.Module's own traversal would walk the registered leaf parameters of the wrapped model instead.
    def parameters(self, recurse=True):
        if getattr(self.module, '_fast', None) is not None:
            return self.module.parameters(recurse)
        return super().parameters(recurse)

    def named_parameters(self, prefix='', recurse=True, remove_duplicate=True):
        if getattr(self.module, '_fast', None) is not None:
            return self.module.named_parameters(prefix + ('.' if prefix else '') + 'module', recurse, remove_duplicate)
        return super().named_parameters(prefix, recurse, remove_duplicate)

    def clone(self, first_order=None, allow_unused=None, allow_nograd=None):
        if first_order is not None and not first_order:
            raise NotImplementedError('only first_order=True is implemented')
        return MAML(self.module.clone_as_learner(), lr=self.lr, first_order=True, allow_unused=self.allow_unused,
                    allow_nograd=self.allow_nograd)

    def adapt(self, loss, first_order=None, allow_unused=None, allow_nograd=None):
        fast = getattr(self.module, '_fast', None)
        if fast is None:
            raise RuntimeError('adapt() must be called on a learner returned by clone()')
        (g,) = torch.autograd.grad(loss, [fast], retain_graph=False, create_graph=False, allow_unused=False)
        self.module.sgd_step(g, self.lr)
