"""SFMLearner with the reference's call surface (models/base_model.py:28-124) whose loss loop runs in
the fused B200 kernels.  The two CNNs are whatever the caller plugs in (Chainer links in the
reference, torch modules in this image); they are outside the hot path and untouched."""
from . import device as D
from .functions import ViewSynthesisLoss
from .lib import LOSS_KEYS


def parse_dict(dic, key, value=None):
    return value if dic is None or key not in dic else dic[key]


class SFMLearner(object):
    """SFMLearner(config, pretrained_model=None, disp_net=..., pose_net=...)

    config keys as in experiments/*.yml `architecture:` -- seq_len, smooth_reg, exp_reg, ssim_rate
    (base_model.py:32-39).  __call__(tgt_img, src_imgs, intrinsics, inv_intrinsics) returns the total
    loss and reports total/pixel/smooth/exp/ssim_loss (base_model.py:119-123) through `reporter`
    (default: stored in `self.last_report`; pass `chainer.report`-style callable to forward them).
    """

    def __init__(self, config, pretrained_model=None, disp_net=None, pose_net=None, reporter=None,
                 B_global=None, raw_disp_scales=0, raw_pose=False):
        self.n_sources = config['seq_len'] - 1
        self.smooth_reg = config['smooth_reg']
        self.exp_reg = config['exp_reg']
        self.ssim_rate = parse_dict(config, 'ssim_rate', 0.0)
        self.disp_net = disp_net
        self.pose_net = pose_net
        self.reporter = reporter
        self.last_report = {}
        self.last_grads = None
        # raw_disp_scales / raw_pose: the nets stop one op early and the kernels apply the disparity activation
        # (disp_net.py:104) / the 0.01 * spatial mean (pose_net.py:52) themselves (ViewSynthesisLoss docstring)
        self.loss_op = ViewSynthesisLoss(self.smooth_reg, self.exp_reg, self.ssim_rate, B_global=B_global,
                                         raw_disp_scales=raw_disp_scales, raw_pose=raw_pose,
                                         edge_aware_smooth=bool(parse_dict(config, 'edge_aware_smooth', False)))

    def __call__(self, tgt_img, src_imgs, intrinsics, inv_intrinsics=None):
        batchsize, n_sources, _, H, W = src_imgs.shape
        stacked_src_imgs = src_imgs.reshape(batchsize, -1, H, W)
        pred_disps = self.disp_net(tgt_img)
        do_exp = self.exp_reg is not None and self.exp_reg > 0
        pred_poses, pred_maskes = self.pose_net(tgt_img, stacked_src_imgs, do_exp=do_exp)
        first = pred_disps[0]
        needs_grad = D.is_torch(first) and any(getattr(t, 'requires_grad', False)
                                               for t in list(pred_disps) + list(pred_poses if isinstance(pred_poses, (tuple, list)) else [pred_poses])
                                               + list(pred_maskes or []))
        if needs_grad:
            import torch
            if not torch.is_grad_enabled():
                needs_grad = False
        if needs_grad:
            from .torch_adapter import view_synthesis_loss
            total, losses = view_synthesis_loss(self.loss_op, tgt_img, src_imgs, intrinsics, pred_disps,
                                                pred_poses, pred_maskes)
        elif type(first).__module__.startswith('chainer'):
            from .chainer_adapter import view_synthesis_loss
            total, losses = view_synthesis_loss(self.loss_op, tgt_img, src_imgs, intrinsics, pred_disps,
                                                pred_poses, pred_maskes)
        else:
            # array-level call (also torch tensors under torch.no_grad(), e.g. validation): PoseNet returns a tuple of
            # S (B, 6) vectors (pose_net.py:52-54); the operator takes one (B, S, 6) array
            if isinstance(pred_poses, (tuple, list)):
                if D.is_torch(first):
                    import torch
                    pred_poses = torch.stack([p.detach() for p in pred_poses], dim=1).contiguous()
                elif D.is_cupy(first):
                    import cupy
                    pred_poses = cupy.ascontiguousarray(cupy.stack(list(pred_poses), axis=1))
                else:
                    raise TypeError('array-level call needs pred_poses as one (B,S,6) device array')
            if D.is_torch(first):
                pred_disps = [t.detach().contiguous() for t in pred_disps]
                pred_maskes = [t.detach().contiguous() for t in pred_maskes] if pred_maskes is not None else None
                pred_poses = pred_poses.detach()
            losses, self.last_grads = self.loss_op.forward_backward(tgt_img, src_imgs, intrinsics, pred_disps,
                                                                    pred_poses, pred_maskes)
            total = losses[0]
        self.last_report = dict(zip(LOSS_KEYS, (losses[i] for i in range(5))))
        if self.reporter is not None:
            for k in LOSS_KEYS:
                self.reporter({k: self.last_report[k]}, self)
        return total
