"""BASELINE config 3: full train step (stand-in DispNet + PoseNet in torch/cuDNN, fused loss kernels, backward,
Adam) at B=4, 128x416 -- step time with the plain seam and with the producer-side fusion, and the share of the
loss path.  Development / reporting aid (the CNNs are not the product)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sfm_learner_chainer_b200 import SFMLearner
from sfm_learner_chainer_b200.synthetic import make_snippets
from tests.standin_nets import DispNetStandIn, PoseNetStandIn


def run(flags, S, raw, B=4, H=128, W=416, steps=30):
    torch.manual_seed(0)
    d = make_snippets(B, S, H, W, seed=0)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tgt, src, K = dev(d['tgt']), dev(d['src']), dev(d['intrinsics'])
    dn, pn = DispNetStandIn(raw_seam=raw).cuda(), PoseNetStandIn(S, raw_seam=raw).cuda()
    model = SFMLearner(dict(seq_len=S + 1, **flags), None, disp_net=dn, pose_net=pn, raw_disp_scales=1 if raw else 0, raw_pose=raw)
    opt = torch.optim.Adam(list(dn.parameters()) + list(pn.parameters()), lr=2e-4)
    def step():
        loss = model(tgt, src, K, K)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # the loss path alone on the same shapes
    with torch.no_grad():
        disps = [x.contiguous() for x in dn(tgt)]
        poses, masks = pn(tgt, src.reshape(B, -1, H, W), do_exp=flags['exp_reg'] > 0)
        poses = poses if raw else torch.stack(list(poses), 1).contiguous()
    op = model.loss_op
    for _ in range(5): op.forward_backward(tgt, src, K, disps, poses, masks)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(200): op.forward_backward(tgt, src, K, disps, poses, masks)
    e1.record(); torch.cuda.synchronize()
    loss_ms = e0.elapsed_time(e1) / 200
    return dict(raw_seam=raw, step_ms=round(ms, 3), loss_path_ms=round(loss_ms, 4), loss_share=round(loss_ms / ms, 4), loss=float(loss))


if __name__ == '__main__':
    flags = dict(smooth_reg=0.0, exp_reg=0.0, ssim_rate=0.0)     # experiments/sfm_learner_v1.yml
    for raw in (False, True):
        print(json.dumps(dict(cfg='cfg3', **run(flags, 2, raw))))
