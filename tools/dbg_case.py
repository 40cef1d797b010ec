import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import sfm_oracle as O
from sfm_learner_chainer_b200 import ViewSynthesisLoss
from sfm_learner_chainer_b200.synthetic import make_snippets
B,S,H,W,seed,harsh = [int(a) for a in sys.argv[1:7]]
flags = dict(smooth_reg=0.1, exp_reg=0.0, ssim_rate=0.15)
d = make_snippets(B,S,H,W,seed=seed,harsh=bool(harsh),rough_disp=bool(seed&1))
L,G,_ = O.sfm_loss(d['tgt'],d['src'],d['intrinsics'],d['disps'],d['poses'],d['logits'],O.LossConfig(**flags))
dev=lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
g=dict(tgt=dev(d['tgt']),src=dev(d['src']),K=dev(d['intrinsics']),disps=[dev(x) for x in d['disps']],poses=dev(d['poses']))
op=ViewSynthesisLoss(**flags)
print('oracle', O.losses_vec(L))
for k in range(2):
    l1=op.forward(g['tgt'],g['src'],g['K'],g['disps'],g['poses'])
    print('fwd   ', l1.cpu().numpy())
    l2,gr=op.forward_backward(g['tgt'],g['src'],g['K'],g['disps'],g['poses'])
    print('fused ', l2.cpu().numpy())
for s in range(4):
    a=gr['gdisps'][s].cpu().numpy(); b_=G['gdisp'][s]
    print('gdisp',s, float(np.abs(a-b_).max()), float(np.abs(b_).max()))
