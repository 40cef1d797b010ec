"""T0: the numpy oracle against fixtures produced by the reference's own code.

tests/golden/*.npz were written by tests/golden/make_golden.py, which executes
/root/reference/models/{base_model,transform,spational_transformer_sampler_interp}.py
unmodified under oracle/chainer_shim.  Nothing here reads /root/reference.
"""
import os

import numpy as np
import pytest

from oracle import sfm_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
CASES = ['v1', 'v1_ssim', 'v1_odom', 'v1_ssim_harsh', 'v1_odom_harsh']


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def _run(g, dt):
    sm, ex, ss = [float(v) for v in g['flags']]
    cfg = O.LossConfig(smooth_reg=sm, exp_reg=ex, ssim_rate=ss)
    c = lambda a: a.astype(dt)
    disps = [c(g['disp%d' % s]) for s in range(4)]
    logits = [c(g['logits%d' % s]) for s in range(4)]
    return O.sfm_loss(c(g['tgt']), c(g['src']), c(g['intrinsics']), disps, c(g['poses']), logits, cfg)


@pytest.mark.parametrize('name', CASES)
def test_oracle_f64_matches_reference_code(name):
    """float64 oracle == float64 run of the reference source: formulas and the analytic backward."""
    g = np.load(os.path.join(GOLD, 'loss_%s.npz' % name))
    L, G, _ = _run(g, np.float64)
    np.testing.assert_allclose(O.losses_vec(L), g['losses_f64'], rtol=1e-12, atol=0)
    assert _rel(G['gpose'], g['gpose_f64']) < 1e-11
    for s in range(4):
        assert _rel(G['gdisp'][s], g['gdisp%d_f64' % s]) < 1e-11
        if G['glogits'] is not None:
            assert _rel(G['glogits'][s], g['glogits%d_f64' % s]) < 1e-11
    assert (G['glogits'] is not None) == (('glogits0_f64') in g.files)


@pytest.mark.parametrize('name', CASES)
def test_oracle_f32_close_to_reference_code(name):
    """canonical fp32 oracle vs the reference run in fp32 and fp64: forward rtol 1e-5 (north star);
    gradients only loosely (fp32 floor flips move individual pixels)."""
    g = np.load(os.path.join(GOLD, 'loss_%s.npz' % name))
    L, G, _ = _run(g, np.float32)
    np.testing.assert_allclose(O.losses_vec(L), g['losses_f64'], rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(O.losses_vec(L), g['losses_f32'], rtol=1e-5, atol=1e-12)
    assert _rel(G['gpose'], g['gpose_f64']) < 5e-3
    for s in range(4):
        assert _rel(G['gdisp'][s], g['gdisp%d_f64' % s]) < 5e-3


@pytest.mark.parametrize('name', ['v1', 'v1_odom_harsh'])
@pytest.mark.parametrize('scale', [0, 2])
def test_warp_stage_matches_reference_code(name, scale):
    """projective_inverse_warp / cam2pixel (transform.py:111-193) stage outputs, float64."""
    g = np.load(os.path.join(GOLD, 'loss_%s.npz' % name))
    B, S, _, H, W = g['src'].shape
    i = S - 1
    h, w = H >> scale, W >> scale
    img = g['warp_s%d_img_f64' % scale]
    stacked = g['src'].astype(np.float64).reshape(B, 3 * S, H, W)
    np.testing.assert_allclose(O.resize_images(stacked, (h, w))[:, 3 * i:3 * i + 3], img, rtol=0, atol=1e-15)
    depth = (1.0 / g['disp%d' % scale].astype(np.float64)).reshape(B, h * w)
    K = g['intrinsics'][:, scale].astype(np.float64)
    pose = g['poses'][:, i].astype(np.float64)
    P, rec = O.projective_inverse_warp(img, depth, pose, K)
    np.testing.assert_allclose(rec['proj'], g['warp_s%d_proj_f64' % scale], rtol=1e-13, atol=1e-13)
    grid = np.stack([rec['grid']['xn'], rec['grid']['yn']], 1).reshape(B, 2, h, w)
    np.testing.assert_allclose(grid, g['warp_s%d_grid_f64' % scale], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(P, g['warp_s%d_P_f64' % scale], rtol=0, atol=1e-10)


def test_interp_sampler_matches_reference_code():
    """models/spational_transformer_sampler_interp.py forward/backward (pure numpy in the reference)."""
    g = np.load(os.path.join(GOLD, 'interp_sampler.npz'))
    y = O.sampler_interp_forward(g['x'], g['grid'])
    np.testing.assert_array_equal(y, g['y'])
    gx, ggrid = O.sampler_interp_backward(g['x'], g['grid'], g['gy'])
    np.testing.assert_array_equal(gx, g['gx'])
    np.testing.assert_allclose(ggrid, g['ggrid'], rtol=1e-6, atol=1e-6)


def test_shard_sum_equals_full_batch():
    """Every F.mean divides by the global batch: partial losses of snippet shards add up."""
    from sfm_learner_chainer_b200.synthetic import make_snippets
    d = make_snippets(4, 2, 32, 104, seed=5)
    cfg = O.LossConfig(smooth_reg=0.1, exp_reg=0.2)
    Lf, Gf, _ = O.sfm_loss(d['tgt'], d['src'], d['intrinsics'], d['disps'], d['poses'], d['logits'], cfg)
    acc = np.zeros(5)
    cfg2 = O.LossConfig(smooth_reg=0.1, exp_reg=0.2, B_global=4)
    for lo in (0, 2):
        sl = slice(lo, lo + 2)
        L, G, _ = O.sfm_loss(d['tgt'][sl], d['src'][sl], d['intrinsics'][sl], [x[sl] for x in d['disps']],
                             d['poses'][sl], [x[sl] for x in d['logits']], cfg2)
        acc += O.losses_vec(L)
        np.testing.assert_allclose(G['gpose'], Gf['gpose'][sl], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(G['gdisp'][0], Gf['gdisp'][0][sl], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(acc, O.losses_vec(Lf), rtol=1e-12)


def test_identity_pose_reproduces_source():
    """T4 property: zero pose => warp is the identity on interior pixels."""
    from sfm_learner_chainer_b200.synthetic import make_snippets
    d = make_snippets(1, 1, 32, 104, seed=6)
    K = d['intrinsics'][:, 0]
    depth = (1.0 / d['disps'][0]).reshape(1, -1)
    P, rec = O.projective_inverse_warp(d['src'][:, 0], depth, np.zeros((1, 6), np.float32), K)
    inb = (rec['grid']['inx'] & rec['grid']['iny']).reshape(32, 104)
    assert inb[1:-1, 1:-1].all()
    np.testing.assert_allclose(P[0][:, inb], d['src'][0, 0][:, inb], atol=2e-4)


def _x_translation_case(dt=np.float64, H=24, W=40, tx=0.0123, depth=2.5, fx=30.0):
    rs = np.random.RandomState(5)
    img = rs.uniform(-1, 1, (1, 3, H, W)).astype(dt)
    K = np.array([[[fx, 0, (W - 1) / 2.0], [0, fx, (H - 1) / 2.0], [0, 0, 1]]], dt)
    pose = np.array([[0, 0, 0, tx, 0, 0]], dt)
    return img, K, pose, np.full((1, H * W), depth, dt), fx * tx / depth


def test_pure_x_translation_is_an_analytic_shift():
    """Constant depth d and a pure x translation t: u = x + fx*t/d for every pixel, so the warp is the source
    shifted by a constant sub-pixel amount."""
    img, K, pose, depth, shift = _x_translation_case()
    P, rec = O.projective_inverse_warp(img, depth, pose, K)
    H, W = img.shape[2:]
    k = int(np.floor(shift))
    f = shift - k
    x = np.arange(W)
    inside = (x + shift > 0) & (x + shift < W - 1)            # strictly inside (-1, 1) after normalisation
    ref = np.zeros_like(img)
    xi = x[inside]
    ref[..., xi] = (1 - f) * img[..., xi + k] + f * img[..., np.minimum(xi + k + 1, W - 1)]
    # rows 0 and H-1 sit exactly on yn = -1 / +1: not strictly inside, hence out of view (transform.py:128-131)
    np.testing.assert_allclose(P[:, :, 1:-1][..., inside], ref[:, :, 1:-1][..., inside], rtol=0, atol=1e-8)   # z = q2 + 1e-10 (transform.py:123)
    # (in float64 the 1e-10 of transform.py:123 pulls the last row just inside; in fp32 -- the device test -- it is out)
    assert np.all(P[..., ~inside] == 0) and np.all(P[:, :, 0] == 0)
