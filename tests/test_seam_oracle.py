"""The seam either side of the loss: oracle restatement of the disparity
activation (models/disp_net.py:104) and the pose reduction (models/pose_net.py:52-53) against fixtures made by
tests/golden/make_golden.py, which runs the reference's own base_model.py and PoseNet.pred_pose and the
reference's DISP_SCALING / MIN_DISP constants under oracle/chainer_shim.  Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

from oracle import sfm_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def run_seam(g, dt):
    sm, ex, ss = [float(v) for v in g['flags']]
    cfg = O.LossConfig(smooth_reg=sm, exp_reg=ex, ssim_rate=ss)
    c = lambda a: a.astype(dt)
    return O.sfm_loss_raw(c(g['tgt']), c(g['src']), c(g['intrinsics']), [c(g['raw_disp%d' % s]) for s in range(4)],
                          c(g['raw_pose']), [c(g['logits%d' % s]) for s in range(4)], cfg,
                          raw_disp_scales=0xF, raw_pose=True)


@pytest.mark.parametrize('name', ['seam_ssim', 'seam_odom'])
def test_seam_oracle_f64_matches_reference_code(name):
    g = np.load(os.path.join(GOLD, name + '.npz'))
    L, G, _ = run_seam(g, np.float64)
    np.testing.assert_allclose(O.losses_vec(L), g['losses_f64'], rtol=1e-12, atol=0)
    assert _rel(G['gpose'], g['gxpose_f64']) < 1e-11
    for s in range(4):
        assert _rel(G['gdisp'][s], g['gx%d_f64' % s]) < 1e-11
        if G['glogits'] is not None:
            assert _rel(G['glogits'][s], g['glogits%d_f64' % s]) < 1e-11


@pytest.mark.parametrize('name', ['seam_ssim', 'seam_odom'])
def test_seam_oracle_f32_close_to_reference_code(name):
    g = np.load(os.path.join(GOLD, name + '.npz'))
    L, G, _ = run_seam(g, np.float32)
    np.testing.assert_allclose(O.losses_vec(L), g['losses_f64'], rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(O.losses_vec(L), g['losses_f32'], rtol=1e-5, atol=1e-12)
    assert _rel(G['gpose'], g['gxpose_f64']) < 5e-3
    for s in range(4):
        assert _rel(G['gdisp'][s], g['gx%d_f64' % s]) < 5e-3


def test_disp_activation_range_and_derivative():
    x = np.linspace(-30, 30, 2001).astype(np.float32)
    d, da = O.disp_activation(x)
    assert d.dtype == np.float32 and d.min() >= np.float32(0.01) and d.max() <= np.float32(10.01)
    x64 = x.astype(np.float64)
    d64, da64 = O.disp_activation(x64)
    np.testing.assert_allclose(d64, 10.0 / (1.0 + np.exp(-x64)) + 0.01, rtol=1e-12)
    eps = 1e-6
    fd = (O.disp_activation(x64 + eps)[0] - O.disp_activation(x64 - eps)[0]) / (2 * eps)
    np.testing.assert_allclose(da64, fd, rtol=1e-6, atol=1e-8)      # finite differences of a value near 10
    np.testing.assert_allclose(d, d64, rtol=3e-7, atol=6e-7)   # tanh form cancels for very negative x (as in Chainer)


@pytest.mark.parametrize('hw', [(1, 4), (2, 5), (3, 9), (8, 16)])
def test_pose_from_raw_is_scaled_mean(hw):
    rs = np.random.RandomState(3)
    x = rs.standard_normal((3, 12) + hw).astype(np.float32)
    p = O.pose_from_raw(x, 2)
    assert p.shape == (3, 2, 6) and p.dtype == np.float32
    ref = 0.01 * x.astype(np.float64).reshape(3, 2, 6, -1).mean(-1)
    np.testing.assert_allclose(p, ref, rtol=2e-6, atol=1e-9)
