"""torch.autograd bridge for the fused loss (plumbing only: device memory + graph hook).

Used where the CNNs are torch modules (this image has no Chainer): DispNet/PoseNet stand-ins produce
pred_disps / pred_poses / pred_maskes, the fused kernels produce the loss and -- in the same pass --
its gradients, and backward() hands them to autograd after rescaling by the upstream gradient."""
import torch

from .functions import ViewSynthesisLoss


class _ViewSynthesisLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, op, tgt, src, intrinsics, n_scales, use_exp, *preds):
        disps = [p.contiguous() for p in preds[:n_scales]]
        poses = preds[n_scales].contiguous()
        logits = [p.contiguous() for p in preds[n_scales + 1:]] if use_exp else None
        losses, grads = op.forward_backward(tgt.contiguous(), src.contiguous(), intrinsics.contiguous(),
                                            [d.detach() for d in disps], poses.detach(),
                                            [l.detach() for l in logits] if logits else None)
        ctx.op, ctx.grads, ctx.shape = op, grads, tuple(src.shape)
        ctx.n_scales, ctx.use_exp = n_scales, use_exp
        ctx.consumed = False
        ctx.mark_non_differentiable(losses)
        return losses[0].clone(), losses

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy, _g_losses):
        # The gradients were produced by the forward's fused pass and are rescaled IN PLACE by the upstream gradient;
        # a second backward through the same node (retain_graph=True) would rescale them again, so it is refused.
        if ctx.consumed:
            raise RuntimeError('view_synthesis_loss: backward through the same loss node twice is not supported '
                               '(its gradients are computed in the forward pass and rescaled in place); call the loss again')
        ctx.consumed = True
        B, S, _, H, W = ctx.shape
        g = ctx.grads
        gy = gy.to(torch.float32).reshape(1).contiguous()
        ctx.op.scale_grads(g, gy, B, S, H, W)            # no-op on the device when gy == 1
        out = list(g['gdisps']) + [g['gposes']]
        if ctx.use_exp:
            out += list(g['glogits'])
        return (None, None, None, None, None, None) + tuple(out)


def view_synthesis_loss(op, tgt, src, intrinsics, pred_disps, pred_poses, pred_maskes=None):
    """-> (total_loss scalar tensor with grad_fn, losses (5,) tensor [total, pixel, smooth, exp, ssim]).

    pred_poses: (B,S,6) tensor or a tuple of S (B,6) tensors as PoseNet returns them (pose_net.py:52-54).
    """
    assert isinstance(op, ViewSynthesisLoss)
    if isinstance(pred_poses, (tuple, list)):
        pred_poses = torch.stack(list(pred_poses), dim=1)
    use_exp = op.use_exp
    preds = list(pred_disps) + [pred_poses] + (list(pred_maskes) if use_exp else [])
    return _ViewSynthesisLossFn.apply(op, tgt, src, intrinsics, op.n_scales, use_exp, *preds)
