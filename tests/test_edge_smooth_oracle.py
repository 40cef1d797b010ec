"""Edge-aware smoothness: oracle restatement of compute_disp_smooth
(models/base_model.py:144-155, call site commented out at :78-80) against the fixture produced by the reference's
own method under oracle/chainer_shim (tests/golden/make_golden.py:make_edge_smooth)."""
import os

import numpy as np

from oracle import sfm_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _run(g, dt):
    B, _, H, W = g['tgt'].shape
    tgt = g['tgt'].astype(dt)
    src = np.zeros((B, 1, 3, H, W), dt)
    K = np.tile(np.eye(3, dtype=dt), (B, 4, 1, 1))
    cfg = O.LossConfig(smooth_reg=float(g['smooth_reg']), edge_aware_smooth=True)
    # only the smoothness term is wanted: the photometric part runs on a zero source image and is ignored
    L, G, _ = O.sfm_loss(tgt, src, K, [g['disp%d' % s].astype(dt) for s in range(4)], np.zeros((B, 1, 6), dt), None, cfg)
    return L, G


def test_edge_smooth_f64_matches_reference_method():
    g = np.load(os.path.join(GOLD, 'edge_smooth.npz'))
    L, _ = _run(g, np.float64)
    np.testing.assert_allclose(L['smooth_loss'], g['loss_f64'], rtol=1e-12)
    # gradient of the smoothness term alone: difference of two oracle runs isolates it from the photometric part
    cfg0 = O.LossConfig(smooth_reg=0.0)
    B, _, H, W = g['tgt'].shape
    tgt = g['tgt'].astype(np.float64)
    src = np.zeros((B, 1, 3, H, W))
    K = np.tile(np.eye(3), (B, 4, 1, 1))
    disps = [g['disp%d' % s].astype(np.float64) for s in range(4)]
    _, G1, _ = O.sfm_loss(tgt, src, K, disps, np.zeros((B, 1, 6)), None, O.LossConfig(smooth_reg=float(g['smooth_reg']), edge_aware_smooth=True))
    _, G0, _ = O.sfm_loss(tgt, src, K, disps, np.zeros((B, 1, 6)), None, cfg0)
    for s in range(4):
        np.testing.assert_allclose(G1['gdisp'][s] - G0['gdisp'][s], g['gdisp%d_f64' % s], rtol=1e-9, atol=1e-15)


def test_edge_smooth_f32_close_to_reference_method():
    g = np.load(os.path.join(GOLD, 'edge_smooth.npz'))
    L, _ = _run(g, np.float32)
    np.testing.assert_allclose(L['smooth_loss'], g['loss_f64'], rtol=1e-5)
    np.testing.assert_allclose(L['smooth_loss'], g['loss_f32'], rtol=1e-5)
