"""Old-style chainer.function.Function: forward_cpu / backward_cpu on numpy arrays."""
import contextlib

import numpy as np
import torch

from .variable import Variable, as_variable
from .utils import type_check


@contextlib.contextmanager
def no_backprop_mode():
    with torch.no_grad():
        yield


class Function(object):
    def check_type_forward(self, in_types):
        pass

    def forward_cpu(self, inputs):
        raise NotImplementedError

    def backward_cpu(self, inputs, grad_outputs):
        raise NotImplementedError

    def forward(self, inputs):
        return self.forward_cpu(inputs)

    def backward(self, inputs, grad_outputs):
        return self.backward_cpu(inputs, grad_outputs)

    def __call__(self, *inputs):
        vs = [as_variable(x) for x in inputs]
        arrays = tuple(v.data for v in vs)
        self.check_type_forward(type_check.get_types(arrays))
        fn = self

        class _Bridge(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *ts):
                outs = fn.forward(tuple(t.detach().numpy() for t in ts))
                ctx.n_out = len(outs)
                return tuple(torch.from_numpy(np.ascontiguousarray(o)) for o in outs)

            @staticmethod
            def backward(ctx, *gys):
                gxs = fn.backward(arrays, tuple(g.numpy() for g in gys))
                return tuple(None if g is None else torch.from_numpy(np.ascontiguousarray(g)) for g in gxs)

        outs = _Bridge.apply(*[v._t for v in vs])
        outs = [Variable(o) for o in outs]
        return outs[0] if len(outs) == 1 else tuple(outs)
