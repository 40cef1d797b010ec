"""-m gpu: BASELINE config 3 -- a full train step of sfm_learner_v1.yml (DispNet + PoseNet forward, the fused
loss kernels, backward through both nets, optimiser step) at B=4, 128x416, random-init weights.

The CNNs are torch stand-ins of the reference's layer shapes (tests/standin_nets.py; Chainer is not
installable here) and are NOT the product; what is checked is the loss path underneath them:
  * the loss equals the oracle evaluated on the nets' outputs (rtol 1e-5);
  * the parameter gradients equal J^T . (oracle gradients w.r.t. the nets' outputs) -- the same autograd graph
    driven by the oracle's gradients instead of the kernels' -- within the fp32 gradient bar;
  * with the producer-side fusion (raw disparity map at scale 0 + raw poseout map) the step produces the same
    loss and the same parameter gradients.
"""
import numpy as np
import pytest

from oracle import sfm_oracle as O
from sfm_learner_chainer_b200.synthetic import make_snippets
from tests.gpu_util import to_dev, host

pytestmark = pytest.mark.gpu


def _build(flags, S, raw):
    import torch
    from sfm_learner_chainer_b200 import SFMLearner
    from tests.standin_nets import DispNetStandIn, PoseNetStandIn
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dn, pn = DispNetStandIn(raw_seam=raw).cuda(), PoseNetStandIn(S, raw_seam=raw).cuda()
    model = SFMLearner(dict(seq_len=S + 1, **flags), None, disp_net=dn, pose_net=pn,
                       raw_disp_scales=1 if raw else 0, raw_pose=raw)
    return model, dn, pn


def _flat_grads(mods):
    import torch
    return torch.cat([p.grad.reshape(-1) for m in mods for p in m.parameters() if p.grad is not None])


@pytest.mark.parametrize('flags,S', [(dict(smooth_reg=0.0, exp_reg=0.0, ssim_rate=0.0), 2),       # sfm_learner_v1.yml
                                     (dict(smooth_reg=0.1, exp_reg=0.2, ssim_rate=0.0), 2)])      # + explainability head
def test_full_train_step_matches_oracle_driven_autograd(flags, S):
    import torch
    B, H, W = 4, 128, 416
    d = make_snippets(B, S, H, W, seed=50)
    tgt, src, K = to_dev(d['tgt']), to_dev(d['src']), to_dev(d['intrinsics'])
    model, dn, pn = _build(flags, S, raw=False)
    opt = torch.optim.Adam(list(dn.parameters()) + list(pn.parameters()), lr=2e-4)       # config_utils.py optimizer block
    before = [p.detach().clone() for p in dn.parameters()][:2]
    loss = model(tgt, src, K, K)
    opt.zero_grad()
    loss.backward()
    g_kernel = _flat_grads([dn, pn]).clone()
    opt.step()
    assert any((a != b).any() for a, b in zip(before, [p.detach() for p in dn.parameters()][:2]))
    # ---- oracle on the nets' outputs (the optimiser step changed the weights: rebuild identical nets)
    model2, dn2, pn2 = _build(flags, S, raw=False)
    disps = dn2(tgt)
    poses, masks = pn2(tgt, src.reshape(B, -1, H, W), do_exp=flags['exp_reg'] > 0)
    pose_t = torch.stack(list(poses), 1)
    L, G, _ = O.sfm_loss(d['tgt'], d['src'], d['intrinsics'], [host(x) for x in disps], host(pose_t),
                         [host(m) for m in masks] if masks is not None else None, O.LossConfig(**flags))
    np.testing.assert_allclose(float(loss.detach()), L['total_loss'], rtol=1e-5)
    outs = list(disps) + [pose_t] + (list(masks) if masks is not None else [])
    gouts = [to_dev(g) for g in G['gdisp']] + [to_dev(G['gpose'])] + ([to_dev(g) for g in G['glogits']] if masks is not None else [])
    torch.autograd.backward(outs, gouts)
    g_oracle = _flat_grads([dn2, pn2])
    assert g_oracle.numel() == g_kernel.numel() > 38e6       # 39.9 M parameters, 38.1 M reached without the explainability head
    rel = float(torch.linalg.norm(g_kernel - g_oracle) / torch.linalg.norm(g_oracle))
    assert rel < 2e-4, rel


def test_full_train_step_with_fused_seam_matches_plain_step():
    import torch
    flags, S = dict(smooth_reg=0.1, exp_reg=0.0, ssim_rate=0.15), 2
    B, H, W = 4, 128, 416
    d = make_snippets(B, S, H, W, seed=51)
    tgt, src, K = to_dev(d['tgt']), to_dev(d['src']), to_dev(d['intrinsics'])
    res = []
    for raw in (False, True):
        model, dn, pn = _build(flags, S, raw=raw)
        loss = model(tgt, src, K, K)
        loss.backward()
        res.append((float(loss.detach()), _flat_grads([dn, pn]).clone(), {k: float(v) for k, v in model.last_report.items()}))
    np.testing.assert_allclose(res[1][0], res[0][0], rtol=2e-6)
    for k in res[0][2]:
        np.testing.assert_allclose(res[1][2][k], res[0][2][k], rtol=2e-6, atol=1e-12)
    rel = float(torch.linalg.norm(res[1][1] - res[0][1]) / torch.linalg.norm(res[0][1]))
    assert rel < 1e-4, rel
