"""chainer.functions stand-in (torch-CPU autograd).  Restates the CPU algorithm of each Chainer op
the reference path calls; see ../README.md.  Differentiable through torch autograd."""
import numpy as np
import torch

from .variable import Variable, as_variable, _to_tensor, _pair


def _t(x):
    return _to_tensor(x)


def _unify(ts):
    dt = ts[0].dtype
    for t in ts[1:]:
        dt = torch.promote_types(dt, t.dtype)
    return [t.to(dt) for t in ts]


def relu(x):
    return Variable(torch.relu(_t(x)))


def clip(x, x_min, x_max):
    return Variable(torch.clamp(_t(x), float(x_min), float(x_max)))


def cos(x):
    return Variable(torch.cos(_t(x)))


def sin(x):
    return Variable(torch.sin(_t(x)))


def exp(x):
    return Variable(torch.exp(_t(x)))


def absolute(x):
    return Variable(torch.abs(_t(x)))


def sigmoid(x):
    t = _t(x)
    return Variable(torch.tanh(t * 0.5) * 0.5 + 0.5)


def stack(xs, axis=0):
    return Variable(torch.stack(_unify([_t(x) for x in xs]), dim=axis))


def concat(xs, axis=1):
    return Variable(torch.cat(_unify([_t(x) for x in xs]), dim=axis))


def hstack(xs):
    ts = _unify([_t(x) for x in xs])
    return Variable(torch.cat(ts, dim=1 if ts[0].dim() > 1 else 0))


def dstack(xs):
    ts = _unify([_t(x) for x in xs])
    ts = [t.reshape(t.shape + (1,) * (3 - t.dim())) if t.dim() < 3 else t for t in ts]
    return Variable(torch.cat(ts, dim=2))


def reshape(x, shape):
    return Variable(_t(x).reshape(tuple(shape)))


def broadcast_to(x, shape):
    return Variable(_t(x).expand(tuple(shape)))


def batch_matmul(a, b, transa=False, transb=False):
    ta, tb = _unify([_t(a), _t(b)])
    if ta.dim() == 2:
        ta = ta.unsqueeze(2)
    if tb.dim() == 2:
        tb = tb.unsqueeze(2)
    if transa:
        ta = ta.transpose(1, 2)
    if transb:
        tb = tb.transpose(1, 2)
    return Variable(torch.matmul(ta, tb))


def batch_inv(a):
    return Variable(torch.linalg.inv(_t(a)))


def where(condition, x, y):
    c = _t(condition).bool()
    tx, ty = _unify([_t(x), _t(y)])
    return Variable(torch.where(c, tx, ty))


def mean(x, axis=None, keepdims=False):
    t = _t(x)
    if axis is None:
        return Variable(t.mean())
    return Variable(t.mean(dim=axis, keepdim=keepdims))


def sum(x, axis=None, keepdims=False):
    t = _t(x)
    if axis is None:
        return Variable(t.sum())
    return Variable(t.sum(dim=axis, keepdim=keepdims))


def split_axis(x, indices_or_sections, axis):
    t = _t(x)
    n = t.shape[axis] // indices_or_sections
    return tuple(Variable(s) for s in torch.split(t, n, dim=axis))


def sigmoid_cross_entropy(x, t, normalize=True, reduce='mean'):
    """loss = -(x*(t - (x>=0)) - log1p(exp(-|x|))), ignore label -1."""
    tx = _t(x)
    tt = torch.from_numpy(np.ascontiguousarray(t)).to(tx.dtype)
    ignore = (tt != -1).to(tx.dtype)
    loss = -(ignore * (tx * (tt - (tx >= 0).to(tx.dtype)) - torch.log1p(torch.exp(-torch.abs(tx)))))
    if reduce == 'no':
        return Variable(loss)
    count = ignore.sum().clamp(min=1) if normalize else tx.shape[0]
    return Variable(loss.sum() / count)


def average_pooling_2d(x, ksize, stride=None, pad=0):
    """im2col + mean over the window: zero padding, pad cells counted in the divisor."""
    t = _t(x)
    stride = ksize if stride is None else stride
    return Variable(torch.nn.functional.avg_pool2d(t, ksize, stride, pad, count_include_pad=True))


def resize_images(x, output_shape):
    """chainer/functions/array/resize_images.py (v4): float64 linspace coordinates, indices clipped
    to [0, n-2], float64 weight products cast to x.dtype, 4-tap gather."""
    t = _t(x)
    B, C, H, W = t.shape
    out_H, out_W = output_shape
    u_1d = np.linspace(0, W - 1, num=out_W)
    v_1d = np.linspace(0, H - 1, num=out_H)
    grid = np.meshgrid(u_1d, v_1d)
    u = grid[0].ravel()
    v = grid[1].ravel()
    u0 = np.floor(u).astype(np.int32)
    u0 = u0.clip(0, W - 2)
    u1 = u0 + 1
    v0 = np.floor(v).astype(np.int32)
    v0 = v0.clip(0, H - 2)
    v1 = v0 + 1
    w1 = (u1 - u) * (v1 - v)
    w2 = (u - u0) * (v1 - v)
    w3 = (u1 - u) * (v - v0)
    w4 = (u - u0) * (v - v0)
    npdt = np.float32 if t.dtype == torch.float32 else np.float64
    w1, w2, w3, w4 = [torch.from_numpy(w.astype(npdt)) for w in (w1, w2, w3, w4)]
    u0, u1, v0, v1 = [torch.from_numpy(i.astype(np.int64)) for i in (u0, u1, v0, v1)]
    y = (w1[None, None, :] * t[:, :, v0, u0] +
         w2[None, None, :] * t[:, :, v0, u1] +
         w3[None, None, :] * t[:, :, v1, u0] +
         w4[None, None, :] * t[:, :, v1, u1])
    return Variable(y.reshape(B, C, out_H, out_W))


def spatial_transformer_sampler(x, grid, **kwargs):
    """chainer/functions/array/spatial_transformer_sampler.py CPU `_forward`: normalised
    coordinates -> pixels (align corners), image zero-padded by one pixel, coordinates clipped to
    the padded image, bilinear blend.  Equals cudnnSpatialTfSampler's zero-padding behaviour."""
    tx, tg = _unify([_t(x), _t(grid)])
    B, C, H, W = tx.shape
    _, _, out_H, out_W = tg.shape
    tg = tg.reshape(B, 2, -1)
    u = tg[:, 0]
    v = tg[:, 1]
    x_pad = torch.nn.functional.pad(tx, (1, 1, 1, 1))
    u = (u + 1) * (W - 1) / 2 + 1
    v = (v + 1) * (H - 1) / 2 + 1
    u_clipped = u.clamp(0, W + 1)
    v_clipped = v.clamp(0, H + 1)
    u0 = torch.floor(u_clipped.detach()).clamp(0, W).long()
    u1 = u0 + 1
    v0 = torch.floor(v_clipped.detach()).clamp(0, H).long()
    v1 = v0 + 1
    w1 = (u1 - u_clipped) * (v1 - v_clipped)
    w2 = (u_clipped - u0) * (v1 - v_clipped)
    w3 = (u1 - u_clipped) * (v_clipped - v0)
    w4 = (u_clipped - u0) * (v_clipped - v0)
    bi = torch.arange(B)[:, None]
    # x_pad[b, :, v, u] -> (B, n, C)
    g = lambda vi, ui: x_pad[bi, :, vi, ui]
    y = (w1[:, :, None] * g(v0, u0) + w2[:, :, None] * g(v0, u1)
         + w3[:, :, None] * g(v1, u0) + w4[:, :, None] * g(v1, u1))
    y = y.reshape(B, out_H, out_W, C).permute(0, 3, 1, 2)
    return Variable(y)
