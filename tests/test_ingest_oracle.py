"""The data layer in front of the loss: oracle restatement of load_as_float_norm
(datasets/kitti/kitti_raw_dataset.py:12-14), data_augmentation and get_multi_scale_intrinsics
(datasets/kitti/kitti_raw_transformed.py:23-93) against the fixture tests/golden/ingest_u8.npz, which
make_golden.py produced by running those reference functions themselves (unmodified, global numpy RNG seeded
per snippet) on synthetic uint8 frames.  Nothing here reads /root/reference."""
import os

import numpy as np

from oracle import sfm_oracle as O
from sfm_learner_chainer_b200.functions import draw_augmentation

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def load():
    g = np.load(os.path.join(GOLD, 'ingest_u8.npz'))
    B, n, H, W, _ = g['frames'].shape
    aug = [draw_augmentation(H, W, np.random.RandomState(int(s))) for s in g['seeds']]
    return g, aug


def test_draws_follow_the_reference_order():
    g, aug = load()
    flips = [a['flip'] for a in aug]
    assert any(flips) and not all(flips)                       # both branches of random_flip are in the fixture
    for a in aug:
        assert 32 <= a['out_h'] <= int(32 * 1.15) and 104 <= a['out_w'] <= int(104 * 1.15)
        assert 0 <= a['off_y'] <= a['out_h'] - 32 and 0 <= a['off_x'] <= a['out_w'] - 104


def test_ingest_oracle_matches_reference_code_bit_for_bit():
    g, aug = load()
    tgt, src, Ks = O.ingest_u8(g['frames'], g['K'], aug)
    np.testing.assert_array_equal(tgt, g['tgt'])
    np.testing.assert_array_equal(src, g['src'])
    np.testing.assert_array_equal(Ks, g['intrinsics'])


def test_ingest_oracle_without_augmentation():
    g, _ = load()
    tgt, src, Ks = O.ingest_u8(g['frames'], g['K'], None)
    np.testing.assert_array_equal(tgt, g['plain_tgt'])
    np.testing.assert_array_equal(Ks, g['plain_intrinsics'])
    assert tgt.min() >= -1.0 and tgt.max() <= 1.0
