import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import sfm_oracle as O
from sfm_learner_chainer_b200 import lib as L
from sfm_learner_chainer_b200.synthetic import make_snippets
d = make_snippets(2, 2, 64, 208, seed=33)
flags = dict(smooth_reg=0.1, exp_reg=0.2, ssim_rate=0.0)
Lo, G, _ = O.sfm_loss(d['tgt'], d['src'], d['intrinsics'], d['disps'], d['poses'], d['logits'], O.LossConfig(**flags))
print('oracle', O.losses_vec(Lo))
lib = L.load()
desc = L.SfmDesc(2, 2, 64, 208, 4, 0, 0.1, 0.2, 0.0, 0)
ctx = C.c_void_p()
L.check(lib.sfm_host_ctx_create(C.byref(desc), C.byref(ctx)))
inp, grads = L.SfmInputs(), L.SfmGrads()
inp.tgt, inp.src = d['tgt'].ctypes.data, d['src'].ctypes.data
inp.intrinsics, inp.poses = d['intrinsics'].ctypes.data, d['poses'].ctypes.data
gd = [np.empty_like(x) for x in d['disps']]; gl = [np.empty_like(x) for x in d['logits']]; gp = np.empty_like(d['poses'])
for s in range(4):
    inp.disps[s], inp.logits[s] = d['disps'][s].ctypes.data, d['logits'][s].ctypes.data
    grads.gdisps[s], grads.glogits[s] = gd[s].ctypes.data, gl[s].ctypes.data
grads.gposes = gp.ctypes.data
losses = np.empty(5, np.float32)
for k in range(3):
    L.check(lib.sfm_loss_step_host(ctx, C.byref(inp), losses.ctypes.data, C.byref(grads)))
    print('host path', losses)
