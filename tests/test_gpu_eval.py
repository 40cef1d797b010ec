"""-m gpu: sfm_eval_depth (resize + clip + masked exact-median scaling + seven depth errors on the device) against
the oracle and the fixture made with the reference's own compute_depth_errors."""
import os

import numpy as np
import pytest

from oracle import sfm_oracle as O
from tests.gpu_util import to_dev, host

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('parity', ['even', 'odd'])
def test_eval_matches_reference_fixture(parity):
    from sfm_learner_chainer_b200 import evaluate_depth_batch
    g = np.load(os.path.join(GOLD, 'eval_depth.npz'))
    out = host(evaluate_depth_batch(to_dev(g['pred_depth']), to_dev(g['gt']), to_dev(g['mask_' + parity].astype(np.uint8)),
                                    float(g['min_depth']), float(g['max_depth'])))
    np.testing.assert_array_equal(out[7], np.float32(g['scale_' + parity]))       # exact medians -> identical scale factor
    np.testing.assert_allclose(out[:7], g['errors_' + parity], rtol=1e-5)         # fp64 sums here, fp32 pairwise means in numpy


@pytest.mark.parametrize('B,h,w,Hg,Wg,seed,frac', [(1, 128, 416, 375, 1242, 0, 0.3), (4, 128, 416, 375, 1242, 1, 0.05),
                                                  (2, 32, 104, 32, 104, 2, 1.0), (3, 40, 72, 121, 217, 3, 0.002)])
def test_eval_matches_oracle(B, h, w, Hg, Wg, seed, frac):
    from sfm_learner_chainer_b200 import evaluate_depth_batch
    rs = np.random.RandomState(seed)
    gt = rs.uniform(0.5, 80.0, (B, Hg, Wg)).astype(np.float32)
    pred = rs.uniform(0.0, 3.0, (B, 1, h, w)).astype(np.float32) ** 3
    mask = rs.uniform(0, 1, (B, Hg, Wg)) < frac
    mask[0, 0, 0] = True
    # many equal values around the median: ties must not disturb the selection
    gt[mask & (rs.uniform(0, 1, gt.shape) < 0.3)] = np.float32(20.0)
    err, scale = O.evaluate_depth_batch(pred, gt, mask, 1e-3, 80.0)
    out = host(evaluate_depth_batch(to_dev(pred), to_dev(gt), to_dev(mask.astype(np.uint8)), 1e-3, 80.0))
    np.testing.assert_array_equal(out[7], np.float32(scale))
    np.testing.assert_allclose(out[:7], err, rtol=2e-5)


def test_eval_argument_errors():
    from sfm_learner_chainer_b200 import lib as L
    lib = L.load()
    assert lib.sfm_eval_depth_scratch_bytes(0, 10, 10) == 0
    assert lib.sfm_eval_depth(1, 8, 8, 16, 16, None, None, None, 1e-3, 80.0, None, None, None) == L.SFM_E_NULL_POINTER
    assert lib.sfm_eval_depth(1, 8, 8, 16, 16, None, None, None, 0.0, 80.0, None, None, None) == L.SFM_E_INVALID_DESC
