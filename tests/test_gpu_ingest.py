"""-m gpu: sfm_ingest_u8 (uint8 frames -> normalised, augmented float images + multi-scale intrinsics in one
device gather) against the oracle and against the fixture made by the reference's own dataset code.
Bit-exact: the kernel only selects, converts and blends with individually rounded fp32 operations."""
import os

import numpy as np
import pytest

from oracle import sfm_oracle as O
from sfm_learner_chainer_b200.functions import draw_augmentation
from tests.gpu_util import to_dev, host

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_ingest_matches_reference_fixture_bit_for_bit():
    from sfm_learner_chainer_b200 import ingest_u8
    g = np.load(os.path.join(GOLD, 'ingest_u8.npz'))
    B, n, H, W, _ = g['frames'].shape
    aug = [draw_augmentation(H, W, np.random.RandomState(int(s))) for s in g['seeds']]
    tgt, src, Ks = ingest_u8(to_dev(g['frames']), to_dev(g['K']), aug)
    np.testing.assert_array_equal(host(tgt), g['tgt'])
    np.testing.assert_array_equal(host(src), g['src'])
    np.testing.assert_array_equal(host(Ks), g['intrinsics'])
    tgt, src, Ks = ingest_u8(to_dev(g['frames']), to_dev(g['K']), None)
    np.testing.assert_array_equal(host(tgt), g['plain_tgt'])
    np.testing.assert_array_equal(host(Ks), g['plain_intrinsics'])


@pytest.mark.parametrize('B,S,H,W,seed', [(4, 2, 128, 416, 0), (2, 4, 128, 416, 1), (3, 2, 40, 72, 2), (1, 3, 256, 832, 3)])
def test_ingest_matches_oracle_bit_for_bit(B, S, H, W, seed):
    from sfm_learner_chainer_b200 import ingest_u8
    rs = np.random.RandomState(seed)
    frames = rs.randint(0, 256, (B, 1 + S, H, W, 3)).astype(np.uint8)
    K = np.tile(np.array([[241.67 * W / 416, 0, 204.2 * W / 416], [0, 246.28 * H / 128, 59.0 * H / 128], [0, 0, 1]], np.float32), (B, 1, 1))
    aug = [draw_augmentation(H, W, rs) for _ in range(B)]
    aug[0].update(out_h=H, out_w=W, off_y=0, off_x=0, x_scaling=1.0, y_scaling=1.0)       # identity scaling edge case
    if B > 1:
        aug[1].update(off_y=aug[1]['out_h'] - H, off_x=aug[1]['out_w'] - W, flip=True)        # crop window at the far corner
    t_ref, s_ref, K_ref = O.ingest_u8(frames, K, aug)
    tgt, src, Ks = ingest_u8(to_dev(frames), to_dev(K), aug)
    np.testing.assert_array_equal(host(tgt), t_ref)
    np.testing.assert_array_equal(host(src), s_ref)
    np.testing.assert_array_equal(host(Ks), K_ref)


def test_ingest_feeds_the_loss_path():
    """uint8 frames -> ingest -> fused loss == oracle ingest -> oracle loss (rtol 1e-5)."""
    from sfm_learner_chainer_b200 import ingest_u8, ViewSynthesisLoss
    from sfm_learner_chainer_b200.synthetic import make_snippets
    B, S, H, W = 2, 2, 64, 208
    d = make_snippets(B, S, H, W, seed=70)
    rs = np.random.RandomState(70)
    imgs = np.concatenate([d['tgt'][:, None], d['src']], 1)                                  # (B,1+S,3,H,W) in [-1,1]
    frames = np.clip(np.round((imgs.transpose(0, 1, 3, 4, 2) + 1) * 127.5), 1, 255).astype(np.uint8)
    aug = [draw_augmentation(H, W, rs) for _ in range(B)]
    K = d['intrinsics'][:, 0].copy()
    tgt, src, Ks = ingest_u8(to_dev(frames), to_dev(K), aug)
    flags = dict(smooth_reg=0.1, exp_reg=0.0, ssim_rate=0.15)
    op = ViewSynthesisLoss(**flags)
    losses, grads = op.forward_backward(tgt, src, Ks, [to_dev(x) for x in d['disps']], to_dev(d['poses']), None)
    t_ref, s_ref, K_ref = O.ingest_u8(frames, K, aug)
    L, G, _ = O.sfm_loss(t_ref, s_ref, K_ref, d['disps'], d['poses'], None, O.LossConfig(**flags))
    np.testing.assert_allclose(host(losses), O.losses_vec(L), rtol=1e-5, atol=1e-9)


def test_ingest_argument_errors():
    from sfm_learner_chainer_b200 import ingest_u8
    frames = np.zeros((1, 3, 16, 16, 3), np.uint8)
    K = np.eye(3, dtype=np.float32)[None]
    with pytest.raises(TypeError):
        ingest_u8(frames, K)                                    # host arrays: no CPU fallback
    with pytest.raises(ValueError):
        ingest_u8(to_dev(frames), to_dev(K), [dict(out_h=16, out_w=16, off_y=1, off_x=0, flip=False, x_scaling=1.0, y_scaling=1.0)])


def test_host_u8_entry_point_matches_device_path():
    """sfm_loss_step_host_u8_submit (host uint8 frames in) == ingest_u8 + forward_backward on the device."""
    import ctypes as C
    import torch
    from sfm_learner_chainer_b200 import ingest_u8, ViewSynthesisLoss, lib as L
    from sfm_learner_chainer_b200.synthetic import make_snippets
    B, S, H, W = 2, 2, 64, 208
    d = make_snippets(B, S, H, W, seed=71)
    rs = np.random.RandomState(71)
    frames = rs.randint(0, 256, (B, 1 + S, H, W, 3)).astype(np.uint8)
    K = np.ascontiguousarray(d['intrinsics'][:, 0])
    aug = [draw_augmentation(H, W, rs) for _ in range(B)]
    flags = dict(smooth_reg=0.1, exp_reg=0.2, ssim_rate=0.0)
    tgt, src, Ks = ingest_u8(to_dev(frames), to_dev(K), aug)
    l_dev, g_dev = ViewSynthesisLoss(**flags).forward_backward(tgt, src, Ks, [to_dev(x) for x in d['disps']], to_dev(d['poses']),
                                                               [to_dev(x) for x in d['logits']])
    lib = L.load()
    desc = L.SfmDesc(B, S, H, W, 4, 0, flags['smooth_reg'], flags['exp_reg'], flags['ssim_rate'], 0)
    ctx = C.c_void_p()
    L.check(lib.sfm_host_ctx_create(C.byref(desc), C.byref(ctx)))
    try:
        arr = (L.SfmAugment * B)(*[L.SfmAugment(a['out_h'], a['out_w'], a['off_y'], a['off_x'], int(a['flip']), 0, a['x_scaling'], a['y_scaling']) for a in aug])
        inp, grads = L.SfmInputs(), L.SfmGrads()
        inp.poses = d['poses'].ctypes.data
        gd = [np.empty_like(x) for x in d['disps']]
        gl = [np.empty_like(x) for x in d['logits']]
        gp = np.empty_like(d['poses'])
        for s in range(4):
            inp.disps[s], inp.logits[s] = d['disps'][s].ctypes.data, d['logits'][s].ctypes.data
            grads.gdisps[s], grads.glogits[s] = gd[s].ctypes.data, gl[s].ctypes.data
        grads.gposes = gp.ctypes.data
        losses = np.empty(5, np.float32)
        for _ in range(2):
            L.check(lib.sfm_loss_step_host_u8_submit(ctx, frames.ctypes.data, K.ctypes.data, arr, C.byref(inp), losses.ctypes.data, C.byref(grads)))
            L.check(lib.sfm_loss_step_host_wait(ctx))
    finally:
        lib.sfm_host_ctx_destroy(ctx)
    np.testing.assert_array_equal(losses, host(l_dev))
    np.testing.assert_array_equal(gp, host(g_dev['gposes']))
    for s in range(4):
        np.testing.assert_array_equal(gd[s], host(g_dev['gdisps'][s]))
        np.testing.assert_array_equal(gl[s], host(g_dev['glogits'][s]))
