"""-m gpu: producer-side fusions (SfmDesc.raw_disp_scales / raw_pose_hw).

Three layers of evidence:
  1. the stage kernels (sfm_disp_activation, sfm_pose_reduce) against the oracle's restatement of
     models/disp_net.py:104 and models/pose_net.py:52-53 (a few ulp: tanhf differs between libm and CUDA);
  2. raw-input mode == stage kernels + plain mode, BIT FOR BIT on losses, warp indices and masks, and to
     rounding on the gradients (the fused kernels inline the very same device functions);
  3. raw-input mode against the oracle and against the golden fixtures made by the reference's own code.
"""
import os

import numpy as np
import pytest

from oracle import sfm_oracle as O
from sfm_learner_chainer_b200.synthetic import make_snippets, make_raw_seam
from tests.gpu_util import to_dev, host, assert_grad_close

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')

FLAGSETS = {
    'v1': dict(smooth_reg=0.0, exp_reg=0.0, ssim_rate=0.0),
    'v1_ssim': dict(smooth_reg=0.1, exp_reg=0.0, ssim_rate=0.15),
    'v1_odom': dict(smooth_reg=0.1, exp_reg=0.2, ssim_rate=0.0),
}


def _op(flags, **kw):
    from sfm_learner_chainer_b200 import ViewSynthesisLoss
    return ViewSynthesisLoss(flags['smooth_reg'], flags['exp_reg'], flags['ssim_rate'], **kw)


def _ulp_diff(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def test_disp_activation_stage_vs_oracle():
    from sfm_learner_chainer_b200 import disp_activation
    rs = np.random.RandomState(0)
    x = np.concatenate([rs.standard_normal(100000) * 3, np.linspace(-40, 40, 4001), [0.0, -0.0, 1e-30, -88.0, 88.0]]).astype(np.float32)
    d, da = disp_activation(to_dev(x), want_dact=True)
    d_ref, da_ref = O.disp_activation(x)
    # tanhf of libm and of CUDA agree within 2 ulp of tanh (|tanh| <= 1: 1.2e-7 absolute), i.e. 6e-8 in y and
    # 6e-7 in disp = 10 y + 0.01 -- many ulp of a disparity near MIN_DISP, where Chainer's tanh form cancels
    np.testing.assert_allclose(host(d), d_ref, rtol=0, atol=2.5e-6)     # + the roundings of 10 y and of + 0.01 near 10
    mid = np.abs(x) < 1
    assert _ulp_diff(host(d)[mid], d_ref[mid]).max() <= 8
    np.testing.assert_allclose(host(da), da_ref, rtol=2e-6, atol=1e-6)
    assert host(d).min() >= np.float32(0.01) and host(d).max() <= np.float32(10.01)


@pytest.mark.parametrize('B,S,hw', [(2, 2, (1, 4)), (3, 4, (2, 5)), (1, 2, (3, 9)), (2, 3, (8, 16))])
def test_pose_reduce_stage_vs_oracle(B, S, hw):
    from sfm_learner_chainer_b200 import pose_reduce
    rs = np.random.RandomState(1)
    x = rs.standard_normal((B, 6 * S) + hw).astype(np.float32)
    got = host(pose_reduce(to_dev(x), S))
    ref = O.pose_from_raw(x, S)
    if hw[0] * hw[1] < 8:
        np.testing.assert_array_equal(got, ref)          # sequential fp32 sum: order unambiguous (KITTI shape 1x4)
    assert _ulp_diff(got, ref).max() <= 2


@pytest.mark.parametrize('flagset', ['v1', 'v1_ssim', 'v1_odom'])
@pytest.mark.parametrize('mask', [0x1, 0xF])
def test_raw_mode_equals_stage_plus_plain_mode(flagset, mask):
    """Bit-level equivalence: fusing the activation / pose reduction changes no loss, index or mask."""
    from sfm_learner_chainer_b200 import disp_activation, pose_reduce
    flags = FLAGSETS[flagset]
    B, S, H, W = 2, (4 if flagset == 'v1_odom' else 2), 64, 208
    d = make_snippets(B, S, H, W, seed=11)
    raw_disps, raw_pose = make_raw_seam(d, (1, 4), seed=11)
    tgt, src, K = to_dev(d['tgt']), to_dev(d['src']), to_dev(d['intrinsics'])
    logits = [to_dev(x) for x in d['logits']]
    xs = [to_dev(x) for x in raw_disps]
    xp = to_dev(raw_pose)
    # plain mode on stage outputs
    acts = [disp_activation(x, want_dact=True) for x in xs]
    disps = [acts[s][0] if (mask >> s) & 1 else to_dev(d['disps'][s]) for s in range(4)]
    poses = pose_reduce(xp, S)
    plain = _op(flags)
    l0, dbg0 = plain.forward(tgt, src, K, disps, poses, logits, debug=True)
    l0b, g0 = plain.forward_backward(tgt, src, K, disps, poses, logits)
    # raw mode
    ins = [xs[s] if (mask >> s) & 1 else to_dev(d['disps'][s]) for s in range(4)]
    fused = _op(flags, raw_disp_scales=mask, raw_pose=True)
    l1, dbg1 = fused.forward(tgt, src, K, ins, xp, logits, debug=True)
    l1b, g1 = fused.forward_backward(tgt, src, K, ins, xp, logits)
    # the five scalars are sums of per-task fp32 partials and the raw-input instances may run with another task shape
    # (the source-split SSIM variant is plain-mode only): equal to summation order, i.e. a few ulp
    np.testing.assert_allclose(host(l1), host(l0), rtol=5e-7, atol=0)
    np.testing.assert_allclose(host(l1b), host(l0b), rtol=5e-7, atol=0)
    for s in range(4):
        for k in ('u0', 'v0', 'inb', 'P'):
            np.testing.assert_array_equal(host(dbg1[k][s]), host(dbg0[k][s]), err_msg='%s scale %d' % (k, s))
    for s in range(4):
        ref = host(g0['gdisps'][s]).astype(np.float64)
        if (mask >> s) & 1:
            ref = ref * host(acts[s][1]).astype(np.float64)
        assert_grad_close(host(g1['gdisps'][s]), ref, rtol=2e-6, atol_rel=1e-7, what='gdisp[%d]' % s)
        if flags['exp_reg']:
            np.testing.assert_array_equal(host(g1['glogits'][s]), host(g0['glogits'][s]))
    gp = host(g0['gposes']).astype(np.float64).reshape(B, 6 * S, 1, 1) * (0.01 / 4)
    assert tuple(g1['gposes'].shape) == raw_pose.shape
    assert_grad_close(host(g1['gposes']), np.broadcast_to(gp, raw_pose.shape), rtol=2e-6, atol_rel=1e-7, what='gposes')


@pytest.mark.parametrize('name', ['seam_ssim', 'seam_odom'])
def test_raw_mode_vs_reference_fixture(name):
    """Against the fixture made by the reference's own base_model.py + PoseNet.pred_pose + DispNet constants."""
    g = np.load(os.path.join(GOLD, name + '.npz'))
    sm, ex, ss = [float(v) for v in g['flags']]
    flags = dict(smooth_reg=sm, exp_reg=ex, ssim_rate=ss)
    S = g['src'].shape[1]
    op = _op(flags, raw_disp_scales=0xF, raw_pose=True)
    losses, grads = op.forward_backward(to_dev(g['tgt']), to_dev(g['src']), to_dev(g['intrinsics']),
                                        [to_dev(g['raw_disp%d' % s]) for s in range(4)], to_dev(g['raw_pose']),
                                        [to_dev(g['logits%d' % s]) for s in range(4)])
    np.testing.assert_allclose(host(losses), g['losses_f64'], rtol=1e-5, atol=1e-9)
    # fp32 oracle on the same inputs: the gradient bar of the north star (rtol 1e-4)
    cfg = O.LossConfig(**flags)
    L, G, _ = O.sfm_loss_raw(g['tgt'], g['src'], g['intrinsics'], [g['raw_disp%d' % s] for s in range(4)],
                             g['raw_pose'], [g['logits%d' % s] for s in range(4)], cfg, raw_disp_scales=0xF, raw_pose=True)
    np.testing.assert_allclose(host(losses), O.losses_vec(L), rtol=1e-5, atol=1e-9)
    assert_grad_close(host(grads['gposes']), G['gpose'], what='gposes (raw map)')
    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
    assert rel(host(grads['gposes']), g['gxpose_f64']) < 5e-3
    for s in range(4):
        # the disparity differs by a few ulp between CUDA and libm tanhf, which can flip a floor index: compare in norm
        assert rel(host(grads['gdisps'][s]), G['gdisp'][s].astype(np.float64)) < 2e-3, s
        assert rel(host(grads['gdisps'][s]), g['gx%d_f64' % s]) < 5e-3, s
        if ex:
            assert_grad_close(host(grads['glogits'][s]), G['glogits'][s], what='glogits[%d]' % s)


def test_raw_mode_argument_errors():
    import ctypes as C
    from sfm_learner_chainer_b200 import lib as L
    lib = L.load()
    d = L.SfmDesc(2, 2, 64, 208, 4, 0, 0.1, 0.0, 0.15, 0, 0x10, 0)
    assert lib.sfm_workspace_bytes(C.byref(d)) == 0 and b'raw_disp_scales' in lib.sfm_last_error()
    d = L.SfmDesc(2, 2, 64, 208, 4, 0, 0.1, 0.0, 0.15, 0, 0, 129)
    assert lib.sfm_workspace_bytes(C.byref(d)) == 0 and b'raw_pose_hw' in lib.sfm_last_error()
    d = L.SfmDesc(2, 2, 64, 208, 4, 0, 0.1, 0.0, 0.15, L.SFM_FLAG_TABLES_PROVIDED, 0, 4)
    assert lib.sfm_workspace_bytes(C.byref(d)) == 0 and b'TABLES_PROVIDED' in lib.sfm_last_error()
