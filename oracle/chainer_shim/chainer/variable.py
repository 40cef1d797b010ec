import numpy as np
import torch


def _to_tensor(x, like=None):
    if isinstance(x, Variable):
        return x._t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        if x.dtype == np.bool_:
            return torch.from_numpy(np.ascontiguousarray(x))
        return torch.from_numpy(np.ascontiguousarray(x))
    if isinstance(x, (bool, int, float, np.generic)):
        # python / numpy scalars adopt the array dtype (numpy/Chainer behaviour)
        dt = like.dtype if like is not None and like.dtype.is_floating_point else torch.float64
        return torch.tensor(float(x), dtype=dt)
    raise TypeError(type(x))


def _pair(a, b):
    """dtype promotion for mixed fp32 constants / fp64 inputs (exact: the reference's
    hard-coded 'f' constants are zeros and ones)."""
    ta = _to_tensor(a, like=b._t if isinstance(b, Variable) else (b if isinstance(b, torch.Tensor) else None))
    tb = _to_tensor(b, like=ta)
    if isinstance(a, (bool, int, float, np.generic)):
        ta = _to_tensor(a, like=tb)
    if ta.dtype != tb.dtype:
        dt = torch.promote_types(ta.dtype, tb.dtype)
        ta, tb = ta.to(dt), tb.to(dt)
    return ta, tb


class Variable(object):
    def __init__(self, data=None, requires_grad=None):
        if isinstance(data, torch.Tensor):
            self._t = data
        else:
            self._t = torch.from_numpy(np.ascontiguousarray(data))
            if requires_grad is None:
                requires_grad = self._t.dtype.is_floating_point
            if requires_grad:
                self._t.requires_grad_(True)

    # --- array-ish surface
    @property
    def data(self):
        return self._t.detach().numpy()

    array = data

    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def ndim(self):
        return self._t.dim()

    @property
    def grad(self):
        return None if self._t.grad is None else self._t.grad.numpy()

    def backward(self):
        self._t.backward(torch.ones_like(self._t))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return Variable(self._t.reshape(shape))

    def __getitem__(self, idx):
        return Variable(self._t[idx])

    def __len__(self):
        return self._t.shape[0]

    # --- arithmetic
    def __neg__(self):
        return Variable(-self._t)

    def __add__(self, o):
        a, b = _pair(self, o)
        return Variable(a + b)

    __radd__ = __add__

    def __sub__(self, o):
        a, b = _pair(self, o)
        return Variable(a - b)

    def __rsub__(self, o):
        a, b = _pair(o, self)
        return Variable(a - b)

    def __mul__(self, o):
        a, b = _pair(self, o)
        return Variable(a * b)

    __rmul__ = __mul__

    def __imul__(self, o):          # Variable has no in-place mul: makes a new node
        return self.__mul__(o)

    def __truediv__(self, o):
        a, b = _pair(self, o)
        return Variable(a / b)

    def __rtruediv__(self, o):
        a, b = _pair(o, self)
        return Variable(a / b)

    def __pow__(self, p):
        return Variable(self._t ** p)


def as_variable(x):
    return x if isinstance(x, Variable) else Variable(_to_tensor(x))
