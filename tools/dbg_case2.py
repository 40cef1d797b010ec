import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import sfm_oracle as O
from sfm_learner_chainer_b200 import ViewSynthesisLoss
from sfm_learner_chainer_b200.synthetic import make_snippets
B,S,H,W,seed,harsh = 1,2,48,160,1,1
flags = dict(smooth_reg=0.0, exp_reg=0.0, ssim_rate=0.15)
d = make_snippets(B,S,H,W,seed=seed,harsh=bool(harsh),rough_disp=True)
dev=lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
op=ViewSynthesisLoss(**flags)
def run(src, poses):
    L,_,_ = O.sfm_loss(d['tgt'],src,d['intrinsics'],d['disps'],poses,None,O.LossConfig(**flags), want_grads=False)
    l=op.forward(dev(d['tgt']),dev(src),dev(d['intrinsics']),[dev(x) for x in d['disps']],dev(poses))
    return O.losses_vec(L)[[1,4]], l.cpu().numpy()[[1,4]]
print('both     ', *run(d['src'], d['poses']))
print('src0 only', *run(d['src'][:, :1], d['poses'][:, :1]))
print('src1 only', *run(d['src'][:, 1:], d['poses'][:, 1:]))
print('swapped  ', *run(d['src'][:, ::-1], d['poses'][:, ::-1]))
print('src0 x2  ', *run(np.concatenate([d['src'][:, :1]]*2,1), np.concatenate([d['poses'][:, :1]]*2,1)))
