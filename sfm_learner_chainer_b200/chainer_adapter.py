"""Chainer binding of the fused loss: a FunctionNode with the reference's calling conventions
(models/utils.py:33-57 shows the FunctionNode style the reference itself uses: `.apply((x,))[0]`).

Import-guarded: chainer==4.0.0b1 / cupy are not installable in the build image, so this module only
defines the node when `chainer` imports.  INTEGRATION.md shows the three-line change to
models/base_model.py that routes SFMLearner.__call__ through it."""
from .functions import ViewSynthesisLoss

try:                                            # pragma: no cover - exercised only where Chainer exists
    import chainer
    from chainer import function_node
    _HAVE_CHAINER = True
except Exception:                               # noqa: BLE001
    chainer = None
    function_node = None
    _HAVE_CHAINER = False


def make_function_node_class(base):
    """Builds the FunctionNode subclass on top of `base` (chainer.function_node.FunctionNode or a
    duck-typed stand-in in tests)."""

    class ViewSynthesisLossFunction(base):
        """inputs: pred_disps[0..n_scales), pred_poses (B,S,6), [pred_maskes[0..n_scales)]
        output: total_loss (scalar array); the five report values are in `self.losses`."""

        def __init__(self, op, tgt_img, src_imgs, intrinsics):
            super(ViewSynthesisLossFunction, self).__init__()
            self.op = op
            self.tgt_img, self.src_imgs, self.intrinsics = tgt_img, src_imgs, intrinsics
            self.losses = None
            self._grads = None

        def check_type_forward(self, in_types):
            # Same style as spational_transformer_sampler_interp.py:11-24 (type_check.expect on dtype kind, ndim and the
            # shapes that tie the inputs together), so that a wrong net output fails here, before any device work.
            ns, S = self.op.n_scales, int(self.src_imgs.shape[1])
            n_in = ns + 1 + (ns if self.op.use_exp else 0)
            if _HAVE_CHAINER:
                from chainer.utils import type_check
                type_check.expect(in_types.size() == n_in)
                B = int(self.src_imgs.shape[0])
                H, W = int(self.src_imgs.shape[3]), int(self.src_imgs.shape[4])
                for s in range(ns):
                    t = in_types[s]
                    type_check.expect(t.dtype.kind == 'f', t.ndim == 4, t.shape[0] == B, t.shape[1] == 1,
                                      t.shape[2] == (H >> s), t.shape[3] == (W >> s))
                pt = in_types[ns]
                type_check.expect(pt.dtype.kind == 'f', pt.ndim == 3, pt.shape[0] == B, pt.shape[1] == S, pt.shape[2] == 6)
                if self.op.use_exp:
                    for s in range(ns):
                        t = in_types[ns + 1 + s]
                        type_check.expect(t.dtype.kind == 'f', t.ndim == 4, t.shape[0] == B, t.shape[1] == S,
                                          t.shape[2] == (H >> s), t.shape[3] == (W >> s))
            else:
                self._check_inputs_plain(in_types, n_in)

        def _check_inputs_plain(self, inputs, n_in):
            """The same conditions without chainer.utils.type_check (duck-typed bases in tests): raises TypeError."""
            ns, S = self.op.n_scales, int(self.src_imgs.shape[1])
            B, H, W = int(self.src_imgs.shape[0]), int(self.src_imgs.shape[3]), int(self.src_imgs.shape[4])
            if len(inputs) != n_in:
                raise TypeError('expected %d inputs (pred_disps, pred_poses%s), got %d' % (
                    n_in, ', pred_maskes' if self.op.use_exp else '', len(inputs)))
            want = [(B, 1, H >> s, W >> s) for s in range(ns)] + [(B, S, 6)]
            if self.op.use_exp:
                want += [(B, S, H >> s, W >> s) for s in range(ns)]
            for k, (a, shp) in enumerate(zip(inputs, want)):
                if 'float32' not in str(a.dtype):
                    raise TypeError('input %d: expected dtype float32, got %s' % (k, a.dtype))
                if tuple(a.shape) != shp:
                    raise TypeError('input %d: expected shape %s, got %s' % (k, shp, tuple(a.shape)))

        def forward(self, inputs):
            ns = self.op.n_scales
            if not _HAVE_CHAINER:                  # chainer's FunctionNode.apply runs check_type_forward itself
                self._check_inputs_plain(inputs, ns + 1 + (ns if self.op.use_exp else 0))
            disps, poses = list(inputs[:ns]), inputs[ns]
            logits = list(inputs[ns + 1:]) if self.op.use_exp else None
            self.losses, self._grads = self.op.forward_backward(self.tgt_img, self.src_imgs, self.intrinsics,
                                                                disps, poses, logits)
            return self.losses[0:1].reshape(()),

        def backward(self, indexes, grad_outputs):
            gy = grad_outputs[0]
            gy_arr = getattr(gy, 'data', gy)
            B, S, _, H, W = self.src_imgs.shape
            g = self.op.scale_grads(self._grads, gy_arr.reshape(1), B, S, H, W)
            outs = list(g['gdisps']) + [g['gposes']] + (list(g['glogits']) if self.op.use_exp else [])
            wrap = (lambda a: chainer.Variable(a)) if _HAVE_CHAINER else (lambda a: a)
            return tuple(wrap(outs[i]) for i in indexes)

    return ViewSynthesisLossFunction


ViewSynthesisLossFunction = make_function_node_class(function_node.FunctionNode) if _HAVE_CHAINER else None


def view_synthesis_loss(op, tgt_img, src_imgs, intrinsics, pred_disps, pred_poses, pred_maskes=None):
    """Chainer-facing call: returns (total_loss Variable, losses array of 5)."""
    if not _HAVE_CHAINER:
        raise ImportError('chainer is not installed; use torch_adapter or the array-level API')
    import chainer.functions as F
    if isinstance(pred_poses, (tuple, list)):
        pred_poses = F.stack(list(pred_poses), axis=1)
    fn = ViewSynthesisLossFunction(op, tgt_img, src_imgs, intrinsics)
    inputs = tuple(pred_disps) + (pred_poses,) + (tuple(pred_maskes) if op.use_exp else ())
    loss, = fn.apply(inputs)
    return loss, fn.losses
