"""One small call of every kernel of libsfmloss for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
  compute-sanitizer --tool racecheck python tools/sanitize_case.py
Shapes are tiny (the tools slow kernels down 10-100x) but cover the strip tails (widths that are not multiples of 28),
both marching kernels, the source-split SSIM variant, raw-input mode, both smoothness kernels, ingest and evaluation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sfm_learner_chainer_b200 import (ViewSynthesisLoss, ingest_u8, evaluate_depth_batch, disp_activation, pose_reduce,
                                      projective_inverse_warp, projective_inverse_warp_backward, SpatialTransformerSamplerInterp)
from sfm_learner_chainer_b200.functions import draw_augmentation
from sfm_learner_chainer_b200.synthetic import make_snippets, make_raw_seam

dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
rs = np.random.RandomState(0)
for (B, S, H, W) in ((2, 2, 40, 72), (1, 4, 32, 104)):
    d = make_snippets(B, S, H, W, seed=1, harsh=True)
    raw_disps, raw_pose = make_raw_seam(d, (1, 4), seed=1)
    g = dict(tgt=dev(d['tgt']), src=dev(d['src']), K=dev(d['intrinsics']), disps=[dev(x) for x in d['disps']], poses=dev(d['poses']),
             logits=[dev(x) for x in d['logits']])
    for flags in (dict(smooth_reg=0.1, exp_reg=0.0, ssim_rate=0.15), dict(smooth_reg=0.1, exp_reg=0.2, ssim_rate=0.0),
                  dict(smooth_reg=0.0, exp_reg=0.0, ssim_rate=0.0)):
        for nw in ('1', '2'):
            os.environ['SFM_SSIM_NW'] = nw
            op = ViewSynthesisLoss(**flags)
            op.forward(g['tgt'], g['src'], g['K'], g['disps'], g['poses'], g['logits'], debug=True)
            op.forward_backward(g['tgt'], g['src'], g['K'], g['disps'], g['poses'], g['logits'])
        os.environ.pop('SFM_SSIM_NW')
        ViewSynthesisLoss(edge_aware_smooth=True, **flags).forward_backward(g['tgt'], g['src'], g['K'], g['disps'], g['poses'], g['logits'])
        ViewSynthesisLoss(raw_disp_scales=0xF, raw_pose=True, **flags).forward_backward(
            g['tgt'], g['src'], g['K'], [dev(x) for x in raw_disps], dev(raw_pose), g['logits'])
    frames = dev(rs.randint(0, 256, (B, 1 + S, H, W, 3)).astype(np.uint8))
    ingest_u8(frames, dev(d['intrinsics'][:, 0].copy()), [draw_augmentation(H, W, rs) for _ in range(B)])
    imgs = dev(d['src'][:, 0])
    depth = dev((1.0 / d['disps'][0]).reshape(B, H * W))
    out = projective_inverse_warp(imgs, depth, dev(d['poses'][:, 0].copy()), dev(d['intrinsics'][:, 0].copy()))
    projective_inverse_warp_backward(imgs, depth, dev(d['poses'][:, 0].copy()), dev(d['intrinsics'][:, 0].copy()), torch.ones_like(out))
x = dev(rs.uniform(-1, 1, (2, 3, 12, 20)).astype(np.float32))
grid = dev(np.stack([rs.uniform(-3, 22, (2, 9, 17)), rs.uniform(-3, 14, (2, 9, 17))], 1).astype(np.float32))
samp = SpatialTransformerSamplerInterp()
y, = samp.forward_gpu((x, grid))
samp.backward_gpu((x, grid), (torch.ones_like(y),))
disp_activation(dev(rs.standard_normal(1000).astype(np.float32)), want_dact=True)
pose_reduce(dev(rs.standard_normal((2, 12, 2, 5)).astype(np.float32)), 2)
gt = dev(rs.uniform(1, 60, (2, 47, 155)).astype(np.float32))
evaluate_depth_batch(dev(rs.uniform(1, 20, (2, 1, 16, 52)).astype(np.float32)), gt, (gt > 20).to(torch.uint8), 1e-3, 80.0)
# the in-kernel cross-GPU sum with a world of one (peer stores into the rank's own slot array)
from sfm_learner_chainer_b200.distributed import PeerLossSum
peer = PeerLossSum(0, 1)
d = make_snippets(2, 2, 40, 72, seed=2)
op = ViewSynthesisLoss(0.1, 0.0, 0.15)
for _ in range(2):
    op.forward_backward(dev(d['tgt']), dev(d['src']), dev(d['intrinsics']), [dev(x) for x in d['disps']], dev(d['poses']), None, peer=peer)
torch.cuda.synchronize()
peer.close()
torch.cuda.synchronize()
print('sanitize_case: done')
