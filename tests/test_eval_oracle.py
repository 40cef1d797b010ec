"""Inference-side depth evaluation: oracle restatement of evaluate.py:94-103 +
kitti_eval/depth_util.py:6-22 against the fixture made with the reference's own compute_depth_errors."""
import os

import numpy as np
import pytest

from oracle import sfm_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('parity', ['even', 'odd'])
def test_eval_oracle_matches_reference_code(parity):
    g = np.load(os.path.join(GOLD, 'eval_depth.npz'))
    err, scale = O.evaluate_depth_batch(g['pred_depth'], g['gt'], g['mask_' + parity], float(g['min_depth']), float(g['max_depth']))
    assert int(g['mask_' + parity].sum()) % 2 == (0 if parity == 'even' else 1)      # both np.median branches
    np.testing.assert_array_equal(np.float32(scale), np.float32(g['scale_' + parity]))
    np.testing.assert_array_equal(err, g['errors_' + parity])
