"""-m gpu: the CUDA path (through the C ABI) against the numpy oracle and the golden fixtures."""
import os

import numpy as np
import pytest

from oracle import sfm_oracle as O
from sfm_learner_chainer_b200.synthetic import make_snippets
from tests.gpu_util import to_dev, dev_inputs, host, oracle_tables, assert_grad_close

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')

FLAGSETS = {
    'v1': dict(smooth_reg=0.0, exp_reg=0.0, ssim_rate=0.0),        # experiments/sfm_learner_v1.yml
    'v1_ssim': dict(smooth_reg=0.1, exp_reg=0.0, ssim_rate=0.15),  # experiments/sfm_learner_v1_ssim.yml
    'v1_odom': dict(smooth_reg=0.1, exp_reg=0.2, ssim_rate=0.0),   # experiments/sfm_learner_v1_odom.yml
    'exp_ssim_rate': dict(smooth_reg=0.05, exp_reg=0.3, ssim_rate=0.25),  # exp branch wins, (1-ssim_rate) weight stays
}


def _op(flags, **kw):
    from sfm_learner_chainer_b200 import ViewSynthesisLoss
    return ViewSynthesisLoss(flags['smooth_reg'], flags['exp_reg'], flags['ssim_rate'], **kw)


def _oracle(d, flags, **kw):
    cfg = O.LossConfig(**flags)
    return O.sfm_loss(d['tgt'], d['src'], d['intrinsics'], d['disps'], d['poses'], d['logits'], cfg, **kw)


# ---------------------------------------------------------------- T1: integer work, bit exact
@pytest.mark.parametrize('flagset', ['v1', 'v1_ssim'])
@pytest.mark.parametrize('B,S,H,W,seed,harsh', [(2, 2, 32, 104, 0, False), (1, 4, 48, 160, 1, True),
                                                (2, 2, 128, 416, 2, False), (1, 3, 40, 72, 3, True)])
def test_indices_masks_and_warp_bit_exact(flagset, B, S, H, W, seed, harsh):
    d = make_snippets(B, S, H, W, seed=seed, harsh=harsh, rough_disp=bool(seed & 1))
    flags = FLAGSETS[flagset]
    proj, kinv = oracle_tables(O, d)
    L, _, dbg = _oracle(d, flags, want_grads=False, want_debug=True,
                        proj_override=np.concatenate([proj, np.zeros_like(proj[..., :1, :])], -2), kinv_override=kinv)
    g = dev_inputs(d)
    losses, gd = _op(flags).forward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'],
                                    proj=to_dev(proj), kinv=to_dev(kinv), debug=True)
    for s in range(4):
        np.testing.assert_array_equal(host(gd['u0'][s]), dbg['u0'][s], err_msg='u0 scale %d' % s)
        np.testing.assert_array_equal(host(gd['v0'][s]), dbg['v0'][s], err_msg='v0 scale %d' % s)
        np.testing.assert_array_equal(host(gd['inb'][s]).astype(bool), dbg['inb'][s], err_msg='inb scale %d' % s)
        np.testing.assert_array_equal(host(gd['P'][s]), dbg['P'][s], err_msg='warped image scale %d' % s)
    np.testing.assert_allclose(host(losses), O.losses_vec(L), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize('H,W', [(32, 104), (128, 416), (40, 72)])
def test_pyramid_bit_exact(H, W):
    d = make_snippets(2, 2, H, W, seed=4)
    tp, sp = _op(FLAGSETS['v1']).pyramid(to_dev(d['tgt']), to_dev(d['src']))
    for s in range(4):
        h, w = H >> s, W >> s
        np.testing.assert_array_equal(host(tp[s]), O.resize_images(d['tgt'], (h, w)))
        ref = O.resize_images(d['src'].reshape(2, 6, H, W), (h, w)).reshape(2, 2, 3, h, w)
        np.testing.assert_array_equal(host(sp[s]), ref)


@pytest.mark.parametrize('harsh', [False, True])
def test_device_tables_match_host_tables(harsh):
    """device prologue (pose_vec2mat/euler2mat/proj_tgt_to_src/batch_inv) vs the oracle's tables."""
    d = make_snippets(8, 4, 128, 416, seed=5, harsh=harsh)
    d['poses'][0, 0, :3] = [3.5, -4.0, 0.3]          # exercises the clip to [-pi, pi] (transform.py:23)
    proj, kinv = oracle_tables(O, d)
    gp, gk = _op(FLAGSETS['v1']).build_tables(to_dev(d['poses']), to_dev(d['intrinsics']), 128, 416)
    np.testing.assert_array_equal(host(gk), kinv)
    # sin/cos are fp64-evaluated on both sides; allow a last-bit difference but report exactness
    np.testing.assert_allclose(host(gp), proj, rtol=3e-7, atol=1e-9)
    assert np.mean(host(gp) == proj) > 0.999


# ---------------------------------------------------------------- T2/T3: end-to-end forward + backward
@pytest.mark.parametrize('flagset', sorted(FLAGSETS))
@pytest.mark.parametrize('B,S,H,W,seed,harsh', [(2, 2, 32, 104, 10, False), (2, 4, 32, 104, 11, False),
                                                (1, 2, 48, 160, 12, True), (1, 3, 40, 72, 13, True),
                                                (2, 2, 128, 416, 14, False)])
def test_loss_and_gradients_match_oracle(flagset, B, S, H, W, seed, harsh):
    d = make_snippets(B, S, H, W, seed=seed, harsh=harsh, rough_disp=bool(seed & 1))
    flags = FLAGSETS[flagset]
    L, G, _ = _oracle(d, flags)
    g = dev_inputs(d)
    op = _op(flags)
    losses, grads = op.forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    np.testing.assert_allclose(host(losses), O.losses_vec(L), rtol=1e-5, atol=1e-9)       # north star: rtol 1e-5
    assert_grad_close(host(grads['gposes']), G['gpose'], what='gpose')                      # north star: rtol 1e-4
    for s in range(4):
        assert_grad_close(host(grads['gdisps'][s]), G['gdisp'][s], what='gdisp[%d]' % s)
        if G['glogits'] is not None:
            assert_grad_close(host(grads['glogits'][s]), G['glogits'][s], what='glogits[%d]' % s)
    assert (grads['glogits'] is not None) == (G['glogits'] is not None)
    # forward-only and recomputing-backward entry points agree with the fused pass
    l2 = op.forward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    np.testing.assert_allclose(host(l2), host(losses), rtol=1e-6, atol=1e-9)
    gy = to_dev(np.array([2.5], np.float32))
    g2 = op.backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'], gy=gy)
    np.testing.assert_allclose(host(g2['gposes']), 2.5 * host(grads['gposes']), rtol=2e-5, atol=1e-7)
    assert_grad_close(host(g2['gdisps'][0]), 2.5 * host(grads['gdisps'][0]), what='gy scaling')
    ref = [host(x).copy() for x in grads['gdisps']]
    op.scale_grads(grads, gy, B, S, H, W)
    np.testing.assert_allclose(host(grads['gdisps'][1]), 2.5 * ref[1], rtol=1e-6, atol=0)


# BASELINE.json configs[3] (sfm_learner_v1_odom.yml:14-16, 5-frame snippets, 128x416) and configs[4] (256x832, SSIM
# flags of sfm_learner_v1_ssim.yml:14-17) at their own shapes, with the launch variant the full batches pick
# (B = 32: 8 runs per L1 task; B = 64: 128-row strips, one warp per SSIM task, forward records in shared memory) forced
# through the development knobs.
BIG_SHAPES = {
    'cfg4': dict(flagset='v1_odom', B=2, S=4, H=128, W=416, seed=70, env={'SFM_HSEG': '8'}),
    'cfg5': dict(flagset='v1_ssim', B=1, S=2, H=256, W=832, seed=71, env={'SFM_HSEG': '128', 'SFM_SSIM_NW': '1', 'SFM_SSIM_SREC': '1'}),
}


@pytest.mark.parametrize('cfg', sorted(BIG_SHAPES))
def test_loss_and_gradients_match_oracle_at_cfg4_cfg5_shapes(cfg, monkeypatch):
    c = BIG_SHAPES[cfg]
    flags = FLAGSETS[c['flagset']]
    d = make_snippets(c['B'], c['S'], c['H'], c['W'], seed=c['seed'])
    L, G, _ = _oracle(d, flags)
    g = dev_inputs(d)
    for forced in (True, False):                     # the big-batch launch variant, then whatever the policy picks here
        for k, v in c['env'].items():
            monkeypatch.setenv(k, v) if forced else monkeypatch.delenv(k, raising=False)
        losses, grads = _op(flags).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
        np.testing.assert_allclose(host(losses), O.losses_vec(L), rtol=1e-5, atol=1e-9)
        assert_grad_close(host(grads['gposes']), G['gpose'], what='gpose')
        for s in range(4):
            assert_grad_close(host(grads['gdisps'][s]), G['gdisp'][s], what='gdisp[%d]' % s)
            if flags['exp_reg']:
                assert_grad_close(host(grads['glogits'][s]), G['glogits'][s], what='glogits[%d]' % s)


@pytest.mark.parametrize('cfg', sorted(BIG_SHAPES))
def test_indices_masks_and_warp_bit_exact_at_cfg4_cfg5_shapes(cfg):
    c = BIG_SHAPES[cfg]
    flags = FLAGSETS[c['flagset']]
    d = make_snippets(1, c['S'], c['H'], c['W'], seed=c['seed'] + 5, harsh=True)
    proj, kinv = oracle_tables(O, d)
    L, _, dbg = _oracle(d, flags, want_grads=False, want_debug=True,
                        proj_override=np.concatenate([proj, np.zeros_like(proj[..., :1, :])], -2), kinv_override=kinv)
    g = dev_inputs(d)
    losses, gd = _op(flags).forward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'],
                                    proj=to_dev(proj), kinv=to_dev(kinv), debug=True)
    for s in range(4):
        np.testing.assert_array_equal(host(gd['u0'][s]), dbg['u0'][s], err_msg='u0 scale %d' % s)
        np.testing.assert_array_equal(host(gd['v0'][s]), dbg['v0'][s], err_msg='v0 scale %d' % s)
        np.testing.assert_array_equal(host(gd['inb'][s]).astype(bool), dbg['inb'][s], err_msg='inb scale %d' % s)
        np.testing.assert_array_equal(host(gd['P'][s]), dbg['P'][s], err_msg='warped image scale %d' % s)
    np.testing.assert_allclose(host(losses), O.losses_vec(L), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize('name', ['v1', 'v1_ssim', 'v1_odom', 'v1_ssim_harsh', 'v1_odom_harsh'])
def test_against_golden_fixtures(name):
    """CUDA path vs outputs of the reference's own source files (tests/golden/make_golden.py)."""
    gd = np.load(os.path.join(GOLD, 'loss_%s.npz' % name))
    sm, ex, ss = [float(v) for v in gd['flags']]
    op = _op(dict(smooth_reg=sm, exp_reg=ex, ssim_rate=ss))
    disps = [to_dev(gd['disp%d' % s]) for s in range(4)]
    logits = [to_dev(gd['logits%d' % s]) for s in range(4)]
    losses, grads = op.forward_backward(to_dev(gd['tgt']), to_dev(gd['src']), to_dev(gd['intrinsics']), disps,
                                        to_dev(gd['poses']), logits)
    np.testing.assert_allclose(host(losses), gd['losses_f64'], rtol=1e-5, atol=1e-9)
    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
    # the fp64 run of the reference is "true math": fp32 floor flips move single pixels,
    # so this comparison is norm-wise; the element-wise bar is test_loss_and_gradients_match_oracle.
    assert rel(host(grads['gposes']), gd['gpose_f64']) < 5e-3
    for s in range(4):
        assert rel(host(grads['gdisps'][s]), gd['gdisp%d_f64' % s]) < 5e-3
        if ex:
            assert rel(host(grads['glogits'][s]), gd['glogits%d_f64' % s]) < 1e-4
    # Second bar: the fixture's fp32 run of the reference (same fp32 floor decisions up to the op order of the shim's
    # matmuls): element-wise rtol 1e-4 wherever the two fp32 runs took the same side of every floor -- the few pixels
    # whose sampling coordinate crosses an integer between the two are excluded by a robust cut and
    # must stay below 0.5 % of the pixels.
    if 'gdisp0_f32' in gd.files:
        for s in range(4):
            got, want = host(grads['gdisps'][s]).astype(np.float64), gd['gdisp%d_f32' % s].astype(np.float64)
            tol = 1e-4 * np.abs(want) + 5e-5 * np.max(np.abs(want))
            off = np.abs(got - want) > tol
            assert off.mean() < 5e-3, 'scale %d: %.3f %% of the pixels differ from the fp32 reference run' % (s, 100 * off.mean())


# ---------------------------------------------------------------- stage API
@pytest.mark.parametrize('h,w,harsh', [(32, 104, False), (128, 416, True), (9, 7, True)])
def test_projective_inverse_warp_stage(h, w, harsh):
    from sfm_learner_chainer_b200 import projective_inverse_warp, projective_inverse_warp_backward
    rs = np.random.RandomState(3)
    N = 3
    d = make_snippets(N, 1, h * 8, w * 8, seed=21, harsh=harsh)
    imgs = O.resize_images(d['src'][:, 0], (h, w)).astype(np.float32)
    depth = (1.0 / O.resize_images(d['disps'][0], (h, w))).reshape(N, h * w).astype(np.float32)
    K = d['intrinsics'][:, 3]
    pose = d['poses'][:, 0]
    Pref, rec = O.projective_inverse_warp(imgs, depth, pose, K)
    proj = np.ascontiguousarray(rec['proj'][:, :3, :])
    out, u0, v0, inb = projective_inverse_warp(to_dev(imgs), to_dev(depth), to_dev(pose), to_dev(K),
                                               proj=to_dev(proj), kinv=to_dev(rec['Kinv']), return_indices=True)
    np.testing.assert_array_equal(host(u0).reshape(N, -1), rec['taps']['u0'])
    np.testing.assert_array_equal(host(v0).reshape(N, -1), rec['taps']['v0'])
    np.testing.assert_array_equal(host(inb).reshape(N, -1).astype(bool), rec['grid']['inx'] & rec['grid']['iny'])
    np.testing.assert_array_equal(host(out), Pref)
    # the reference passes depth broadcast to 3 rows (base_model.py:81-84)
    out3 = projective_inverse_warp(to_dev(imgs), to_dev(np.broadcast_to(depth[:, None], (N, 3, h * w)).copy()),
                                   to_dev(pose), to_dev(K))
    np.testing.assert_allclose(host(out3), Pref, rtol=0, atol=2e-5)
    # backward against the oracle's chain
    gy = rs.uniform(-1, 1, (N, 3, h, w)).astype(np.float32)
    t64 = {k: (v.astype(np.float64) if v.dtype.kind == 'f' else v) for k, v in rec['taps'].items()}
    gxn, gyn = O.spatial_transformer_sampler_grad(t64, gy.reshape(N, 3, -1).astype(np.float64), h, w)
    gr = rec['grid']
    gxn = gxn * np.where(gr['inx'], 1.0, 2.0)
    gyn = gyn * np.where(gr['iny'], 1.0, 2.0)
    z = gr['z'].astype(np.float64)
    q = gr['q'].astype(np.float64)
    with np.errstate(all='ignore'):
        gq = np.stack([gxn / (z * float(gr['hw'])), gyn / (z * float(gr['hh'])),
                       -(gxn * q[:, 0] / float(gr['hw']) + gyn * q[:, 1] / float(gr['hh'])) / (z * z)], 1)
    gq = np.where(rec['taps']['any_valid'][:, None], gq, 0.0)
    gcam = np.einsum('nkp,nkj->njp', gq, rec['proj'][:, :3, :3].astype(np.float64))
    gdepth_ref = np.sum(gcam * rec['ray'].astype(np.float64), axis=1)
    cam4 = np.concatenate([rec['cam'].astype(np.float64), np.ones((N, 1, h * w))], 1)
    dP = np.einsum('nkp,njp->nkj', gq, cam4)
    gpose_ref, _ = O._pose_backward(pose, [K], [dP])
    gdepth, gposes, gimgs = projective_inverse_warp_backward(to_dev(imgs), to_dev(depth), to_dev(pose), to_dev(K),
                                                             to_dev(gy), want_gimgs=True)
    assert_grad_close(host(gdepth), gdepth_ref, what='gdepth')
    assert_grad_close(host(gposes), gpose_ref, what='gposes')
    # gimgs: adjoint identity <gy, warp(img)> == <gimgs, img> (the warp is linear in the image)
    lhs = float(np.sum(gy.astype(np.float64) * Pref))
    rhs = float(np.sum(host(gimgs).astype(np.float64) * imgs))
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0)


def test_sampler_interp_matches_reference_fixture():
    from sfm_learner_chainer_b200 import SpatialTransformerSamplerInterp, spatial_transformer_sampler_interp
    g = np.load(os.path.join(GOLD, 'interp_sampler.npz'))
    fn = SpatialTransformerSamplerInterp()
    x, grid, gy = to_dev(g['x']), to_dev(g['grid']), to_dev(g['gy'])
    y, = fn.forward_gpu((x, grid))
    np.testing.assert_array_equal(host(y), g['y'])
    np.testing.assert_array_equal(host(spatial_transformer_sampler_interp(x, grid)), g['y'])
    gx, ggrid = fn.backward_gpu((x, grid), (gy,))
    np.testing.assert_array_equal(host(gx), g['gx'])
    np.testing.assert_allclose(host(ggrid), g['ggrid'], rtol=1e-6, atol=1e-6)
    with pytest.raises(TypeError):
        fn.forward_gpu((x, to_dev(g['grid'][:, :1])))          # grid.shape[1] != 2 (interp.py:22)


# ---------------------------------------------------------------- properties at full size
def test_identity_pose_and_all_out_of_view():
    d = make_snippets(2, 2, 128, 416, seed=30)
    d['poses'][:] = 0
    g = dev_inputs(d)
    op = _op(FLAGSETS['v1'])
    _, dbg = op.forward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'], debug=True)
    P = host(dbg['P'][0])
    inb = host(dbg['inb'][0]).astype(bool)
    assert inb[:, :, 1:-1, 1:-1].all()
    m = np.broadcast_to(inb[:, :, None], P.shape)
    np.testing.assert_allclose(P[m], d['src'][m], atol=5e-4)
    # a huge sideways translation throws every pixel out of view: photometric terms are exactly 0
    d['poses'][:, :, 3] = 1e4
    g = dev_inputs(d)
    losses, grads = _op(FLAGSETS['v1_ssim']).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'],
                                                              g['poses'], g['logits'])
    lv = host(losses)
    assert lv[1] == 0.0 and lv[4] == 0.0 and lv[2] > 0
    assert not host(grads['gposes']).any()


def test_pure_x_translation_is_an_analytic_shift():
    """T4 property on the device: constant depth + pure x translation = constant sub-pixel shift of the source."""
    from sfm_learner_chainer_b200 import projective_inverse_warp
    from tests.test_oracle_golden import _x_translation_case
    img, K, pose, depth, shift = _x_translation_case(np.float32)
    out, u0, v0, inb = projective_inverse_warp(to_dev(img), to_dev(depth), to_dev(pose), to_dev(K), return_indices=True)
    P = host(out)
    H, W = img.shape[2:]
    k = int(np.floor(shift))
    f = np.float32(shift - k)
    x = np.arange(W)
    inside = (x + shift > 0.01) & (x + shift < W - 1.01)
    xi = x[inside]
    ref = (1 - f) * img[..., xi + k] + f * img[..., xi + k + 1]
    # rows 0 and H-1 sit exactly on yn = -1 / +1: out of view by the strict rule (transform.py:128-131)
    np.testing.assert_allclose(P[:, :, 1:-1][..., inside], ref[:, :, 1:-1], rtol=0, atol=2e-5)   # fp32 coordinates: f is known to ~1e-5
    np.testing.assert_array_equal(host(u0)[0][1:-1][:, inside], np.broadcast_to(xi + k, (H - 2, xi.size)))
    np.testing.assert_array_equal(host(inb)[0][1:-1][:, inside], 1)
    assert np.all(P[:, :, 0] == 0) and np.all(P[:, :, -1] == 0)


@pytest.mark.parametrize('cfg', ['cfg4', 'cfg2'])
def test_shard_sum_equals_full_batch_at_full_size(cfg):
    """Size-independent property at BASELINE.json's shapes: snippet shards (B_global = full batch) sum to
    the full-batch losses and reproduce its gradients -- the multi-GPU decomposition."""
    from sfm_learner_chainer_b200.synthetic import CONFIGS
    c = dict(CONFIGS[cfg])
    B, S, H, W = c.pop('B'), c.pop('S'), c.pop('H'), c.pop('W')
    B = min(B, 8)
    d = make_snippets(B, S, H, W, seed=31)
    g = dev_inputs(d)
    lf, gf = _op(c).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    acc = np.zeros(5)
    half = B // 2
    for lo in (0, half):
        sl = slice(lo, lo + half)
        ls, gs = _op(c, B_global=B).forward_backward(
            g['tgt'][sl].contiguous(), g['src'][sl].contiguous(), g['intrinsics'][sl].contiguous(),
            [x[sl].contiguous() for x in g['disps']], g['poses'][sl].contiguous(),
            [x[sl].contiguous() for x in g['logits']])
        acc += host(ls).astype(np.float64)
        # the shard may run with another task shape (launch policy), i.e. another fp32 accumulation order per task
        assert_grad_close(host(gs['gposes']), host(gf['gposes'])[sl], what='gposes of shard %d' % lo)
        np.testing.assert_array_equal(host(gs['gdisps'][0]), host(gf['gdisps'][0])[sl])
    np.testing.assert_allclose(acc, host(lf), rtol=2e-6)


def test_batch_permutation_equivariance():
    d = make_snippets(4, 2, 64, 208, seed=32)
    perm = np.array([2, 0, 3, 1])
    flags = FLAGSETS['v1_ssim']
    g = dev_inputs(d)
    l0, g0 = _op(flags).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    dp = dict(tgt=d['tgt'][perm], src=d['src'][perm], intrinsics=d['intrinsics'][perm],
              disps=[x[perm] for x in d['disps']], poses=d['poses'][perm], logits=[x[perm] for x in d['logits']])
    g = dev_inputs(dp)
    l1, g1 = _op(flags).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    np.testing.assert_allclose(host(l1), host(l0), rtol=1e-6)
    np.testing.assert_array_equal(host(g1['gdisps'][0]), host(g0['gdisps'][0])[perm])
    np.testing.assert_allclose(host(g1['gposes']), host(g0['gposes'])[perm], rtol=1e-5, atol=1e-9)


def test_host_buffer_entry_point_matches_device_path():
    import ctypes as C
    from sfm_learner_chainer_b200 import lib as L
    d = make_snippets(2, 2, 64, 208, seed=33)
    flags = FLAGSETS['v1_odom']
    g = dev_inputs(d)
    l_dev, g_dev = _op(flags).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    lib = L.load()
    desc = L.SfmDesc(2, 2, 64, 208, 4, 0, flags['smooth_reg'], flags['exp_reg'], flags['ssim_rate'], 0)
    ctx = C.c_void_p()
    L.check(lib.sfm_host_ctx_create(C.byref(desc), C.byref(ctx)))
    try:
        inp, grads = L.SfmInputs(), L.SfmGrads()
        inp.tgt, inp.src = d['tgt'].ctypes.data, d['src'].ctypes.data
        inp.intrinsics, inp.poses = d['intrinsics'].ctypes.data, d['poses'].ctypes.data
        gd = [np.empty_like(x) for x in d['disps']]
        gl = [np.empty_like(x) for x in d['logits']]
        gp = np.empty_like(d['poses'])
        for s in range(4):
            inp.disps[s], inp.logits[s] = d['disps'][s].ctypes.data, d['logits'][s].ctypes.data
            grads.gdisps[s], grads.glogits[s] = gd[s].ctypes.data, gl[s].ctypes.data
        grads.gposes = gp.ctypes.data
        losses = np.empty(5, np.float32)
        L.check(lib.sfm_loss_step_host(ctx, C.byref(inp), losses.ctypes.data, C.byref(grads)))
    finally:
        lib.sfm_host_ctx_destroy(ctx)
    np.testing.assert_allclose(losses, host(l_dev), rtol=1e-6)
    np.testing.assert_allclose(gp, host(g_dev['gposes']), rtol=1e-5, atol=1e-9)
    for s in range(4):
        np.testing.assert_array_equal(gd[s], host(g_dev['gdisps'][s]))
        np.testing.assert_array_equal(gl[s], host(g_dev['glogits'][s]))


def test_host_submit_wait_with_two_contexts_in_flight():
    """sfm_loss_step_host_submit / _wait: two host contexts used alternately on different inputs reproduce the
    synchronous call bit for bit (each context owns its stream, device buffers and workspace)."""
    import ctypes as C
    import torch
    from sfm_learner_chainer_b200 import lib as L
    lib = L.load()
    flags = FLAGSETS['v1_ssim']
    desc = L.SfmDesc(2, 2, 64, 208, 4, 0, flags['smooth_reg'], flags['exp_reg'], flags['ssim_rate'], 0)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    data = [make_snippets(2, 2, 64, 208, seed=60 + k) for k in range(2)]

    def pack(d):
        h = dict(tgt=pin(d['tgt']), src=pin(d['src']), K=pin(d['intrinsics']), poses=pin(d['poses']),
                 disps=[pin(x) for x in d['disps']], gd=[pin(np.zeros_like(x)) for x in d['disps']],
                 gp=pin(np.zeros_like(d['poses'])), losses=pin(np.zeros(8, np.float32)))
        inp, g = L.SfmInputs(), L.SfmGrads()
        inp.tgt, inp.src, inp.intrinsics, inp.poses = h['tgt'].data_ptr(), h['src'].data_ptr(), h['K'].data_ptr(), h['poses'].data_ptr()
        g.gposes = h['gp'].data_ptr()
        for s in range(4):
            inp.disps[s], g.gdisps[s] = h['disps'][s].data_ptr(), h['gd'][s].data_ptr()
        return h, inp, g

    ctxs = []
    try:
        for _ in range(2):
            ctx = C.c_void_p()
            L.check(lib.sfm_host_ctx_create(C.byref(desc), C.byref(ctx)))
            ctxs.append(ctx)
        ref = []
        for k in range(2):
            h, inp, g = pack(data[k])
            L.check(lib.sfm_loss_step_host(ctxs[0], C.byref(inp), C.c_void_p(h['losses'].data_ptr()), C.byref(g)))
            ref.append((h['losses'].numpy().copy(), [x.numpy().copy() for x in h['gd']], h['gp'].numpy().copy()))
        packs = [pack(data[k]) for k in range(2)]
        for rep in range(6):
            for k in range(2):
                if rep:
                    L.check(lib.sfm_loss_step_host_wait(ctxs[k]))
                h, inp, g = packs[k]
                L.check(lib.sfm_loss_step_host_submit(ctxs[k], C.byref(inp), C.c_void_p(h['losses'].data_ptr()), C.byref(g)))
        for k in range(2):
            L.check(lib.sfm_loss_step_host_wait(ctxs[k]))
            h = packs[k][0]
            np.testing.assert_array_equal(h['losses'].numpy()[:5], ref[k][0][:5])
            np.testing.assert_array_equal(h['gp'].numpy(), ref[k][2])
            for s in range(4):
                np.testing.assert_array_equal(h['gd'][s].numpy(), ref[k][1][s])
    finally:
        for ctx in ctxs:
            lib.sfm_host_ctx_destroy(ctx)
    assert lib.sfm_loss_step_host_wait(None) == L.SFM_E_NULL_POINTER


@pytest.mark.parametrize('flagset', ['v1_ssim', 'v1_odom'])
def test_results_do_not_depend_on_the_task_height(flagset, monkeypatch):
    """The launch policy picks the task height (rows / runs per warp task) from a cost model; any height must give
    the same per-pixel gradients bit for bit (only the order of the fp64 cell sums changes) -- guards the halo and
    segment-boundary logic for heights that are not powers of two."""
    flags = FLAGSETS[flagset]
    d = make_snippets(2, 2, 72, 136, seed=45, harsh=True)
    g = dev_inputs(d)
    op = _op(flags)
    monkeypatch.delenv('SFM_HSEG', raising=False)
    l0, g0 = op.forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    ref = (host(l0), [host(x) for x in g0['gdisps']], host(g0['gposes']))
    for hseg in (4, 5, 8, 9, 11, 13, 17, 23, 37, 64):
        monkeypatch.setenv('SFM_HSEG', str(hseg))
        l1, g1 = op.forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
        np.testing.assert_allclose(host(l1), ref[0], rtol=1e-6, err_msg='hseg %d' % hseg)
        assert_grad_close(host(g1['gposes']), ref[2], what='gposes, hseg %d (fp32 per-task accumulation order)' % hseg)
        for s in range(4):
            np.testing.assert_array_equal(host(g1['gdisps'][s]), ref[1][s], err_msg='hseg %d scale %d' % (hseg, s))


def test_source_split_ssim_kernel_equals_sequential_kernel(monkeypatch):
    """Small two-source batches run the SSIM kernel with the sources split over the two warps of a CTA (NW = 2);
    the per-pixel gradients must equal the one-warp-per-task kernel bit for bit (same fused multiply-adds in the
    same order), losses and pose gradients to rounding of the partial sums."""
    flags = FLAGSETS['v1_ssim']
    for (B, H, W, seed) in ((4, 128, 416, 46), (2, 72, 136, 47)):
        d = make_snippets(B, 2, H, W, seed=seed, harsh=bool(seed & 1))
        g = dev_inputs(d)
        res = []
        for nw in ('1', '2'):
            monkeypatch.setenv('SFM_SSIM_NW', nw)
            l, gr = _op(flags).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
            res.append((host(l), [host(x) for x in gr['gdisps']], host(gr['gposes'])))
        monkeypatch.delenv('SFM_SSIM_NW')
        np.testing.assert_allclose(res[1][0], res[0][0], rtol=1e-6)
        assert_grad_close(res[1][2], res[0][2], what='gposes')
        for s in range(4):
            np.testing.assert_array_equal(res[1][1][s], res[0][1][s], err_msg='gdisp scale %d' % s)
        # and without the smoothness term (the kernel then writes gdisp without reading it)
        flags0 = dict(flags, smooth_reg=0.0)
        out = []
        for nw in ('1', '2'):
            monkeypatch.setenv('SFM_SSIM_NW', nw)
            l, gr = _op(flags0).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
            out.append([host(x) for x in gr['gdisps']])
        monkeypatch.delenv('SFM_SSIM_NW')
        for s in range(4):
            np.testing.assert_array_equal(out[1][s], out[0][s], err_msg='gdisp scale %d (no smoothness)' % s)


def test_chainer_function_node_adapter_with_a_duck_typed_base():
    """the FunctionNode adapter (chainer_adapter.py) on a minimal stand-in for
    chainer.function_node.FunctionNode -- apply() -> forward(inputs) -> one output; backward(indexes, grad_outputs)
    returns the gradients of the requested inputs -- with torch CUDA tensors as the device arrays."""
    import torch
    from sfm_learner_chainer_b200.chainer_adapter import make_function_node_class

    class FakeFunctionNode(object):                      # the two methods of the FunctionNode protocol the adapter relies on
        def apply(self, inputs):
            self.inputs = tuple(inputs)
            return self.forward(self.inputs)

    d = make_snippets(2, 2, 32, 104, seed=35)
    flags = FLAGSETS['v1_odom']
    L, G, _ = _oracle(d, flags)
    g = dev_inputs(d)
    Node = make_function_node_class(FakeFunctionNode)
    fn = Node(_op(flags), g['tgt'], g['src'], g['intrinsics'])
    loss, = fn.apply(tuple(g['disps']) + (g['poses'],) + tuple(g['logits']))
    assert tuple(loss.shape) == ()
    np.testing.assert_allclose(float(loss), L['total_loss'], rtol=1e-5)
    np.testing.assert_allclose(host(fn.losses), O.losses_vec(L), rtol=1e-5, atol=1e-9)     # the five chainer.report values
    gy = torch.full((), 2.0, device='cuda')
    want = (0, 4, 6)                                     # pred_disps[0], pred_poses, pred_maskes[1]
    got = fn.backward(want, (gy,))
    assert len(got) == 3
    assert_grad_close(host(got[0]), 2.0 * G['gdisp'][0], what='gdisp[0]')
    assert_grad_close(host(got[1]), 2.0 * G['gpose'], what='gpose')
    assert_grad_close(host(got[2]), 2.0 * G['glogits'][1], what='glogits[1]')


@pytest.mark.parametrize('flagset', ['v1_ssim', 'v1_odom'])
def test_reuse_pyramid_flag(flagset):
    """sfm_pyramid + SFM_FLAG_REUSE_PYRAMID (pyramid built ahead, e.g. for the next batch on a side stream) gives the
    same result as the all-in-one call, bit for bit."""
    flags = FLAGSETS[flagset]
    d = make_snippets(2, 2, 64, 208, seed=36)
    g = dev_inputs(d)
    args = (g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    l0, g0 = _op(flags).forward_backward(*args)
    op = _op(flags)
    op.build_pyramid(g['tgt'], g['src'])
    l1, g1 = op.forward_backward(*args, reuse_pyramid=True)
    np.testing.assert_array_equal(host(l1), host(l0))
    np.testing.assert_array_equal(host(g1['gposes']), host(g0['gposes']))
    for s in range(4):
        np.testing.assert_array_equal(host(g1['gdisps'][s]), host(g0['gdisps'][s]))


def test_pyramid_built_on_a_side_stream_is_ordered_by_the_operator():
    """build_pyramid on a side stream, the reusing loss call on the main stream, alternating batches through ONE
    operator (one workspace): the operator's own events must order pyramid -> loss -> next pyramid, so every
    result equals the all-in-one call bit for bit."""
    import torch
    flags = FLAGSETS['v1_ssim']
    sets = [dev_inputs(make_snippets(4, 2, 128, 416, seed=80 + k)) for k in range(2)]
    ref = []
    for g in sets:
        l, gr = _op(flags).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
        ref.append((host(l), [host(x) for x in gr['gdisps']], host(gr['gposes'])))
    op = _op(flags)
    side = torch.cuda.Stream()
    outs = []
    torch.cuda.synchronize()
    for it in range(12):
        g = sets[it & 1]
        with torch.cuda.stream(side):
            op.build_pyramid(g['tgt'], g['src'])
        outs.append(op.forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'],
                                        reuse_pyramid=True))
    torch.cuda.synchronize()
    for it, (l, gr) in enumerate(outs):
        rl, rg, rp = ref[it & 1]
        np.testing.assert_array_equal(host(l), rl)
        np.testing.assert_array_equal(host(gr['gposes']), rp)
        for s in range(4):
            np.testing.assert_array_equal(host(gr['gdisps'][s]), rg[s])


@pytest.mark.parametrize('flagset', ['v1', 'v1_ssim', 'v1_odom'])
@pytest.mark.parametrize('B,S,H,W', [(1, 1, 32, 32),      # one source view, smallest legal size (4x4 at the coarsest scale)
                                     (3, 3, 36, 60),      # odd source count, sizes that are not multiples of 8 (4x7 at scale 3)
                                     (1, 5, 32, 40),      # five sources: three passes of the two-source kernel
                                     (2, 2, 33, 57)])     # odd sizes: H >> s and W >> s truncate
def test_edge_shapes_match_oracle(flagset, B, S, H, W):
    flags = FLAGSETS[flagset]
    d = make_snippets(B, S, H, W, seed=48, harsh=True)
    L, G, _ = _oracle(d, flags)
    g = dev_inputs(d)
    losses, grads = _op(flags).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    np.testing.assert_allclose(host(losses), O.losses_vec(L), rtol=1e-5, atol=1e-9)
    assert_grad_close(host(grads['gposes']), G['gpose'], what='gposes')
    for s in range(4):
        assert_grad_close(host(grads['gdisps'][s]), G['gdisp'][s], what='gdisp[%d]' % s)
        if flags['exp_reg']:
            assert_grad_close(host(grads['glogits'][s]), G['glogits'][s], what='glogits[%d]' % s)


def test_shapes_below_the_minimum_are_rejected():
    from sfm_learner_chainer_b200 import lib as L
    d = make_snippets(1, 2, 24, 40, seed=49)                  # 3x5 at the coarsest scale: below 4x4
    g = dev_inputs(d)
    with pytest.raises(L.SfmError) as e:
        _op(FLAGSETS['v1']).forward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    assert e.value.code == L.SFM_E_INVALID_SHAPE or e.value.code == L.SFM_E_INVALID_DESC


def test_torch_autograd_bridge_and_model_surface():
    """SFMLearner.__call__ surface (base_model.py:48-124) with stub nets: loss, five reports, gradients
    reaching the producers of pred_disps / pred_poses / pred_maskes."""
    import torch
    from sfm_learner_chainer_b200 import SFMLearner
    d = make_snippets(2, 2, 32, 104, seed=34)
    flags = FLAGSETS['v1_odom']
    L, G, _ = _oracle(d, flags)
    g = dev_inputs(d)
    disps = [x.clone().requires_grad_(True) for x in g['disps']]
    poses = [g['poses'][:, i].clone().requires_grad_(True) for i in range(2)]
    masks = [x.clone().requires_grad_(True) for x in g['logits']]
    reports = {}
    model = SFMLearner(dict(seq_len=3, **flags), None, disp_net=lambda t: disps,
                       pose_net=lambda t, s, do_exp=True: (tuple(poses), masks if do_exp else None),
                       reporter=lambda kv, obs: reports.update(kv))
    loss = model(g['tgt'], g['src'], g['intrinsics'], g['intrinsics'])
    (3.0 * loss).backward()
    assert sorted(reports) == sorted(O.LOSS_KEYS)
    np.testing.assert_allclose(float(loss.detach()), L['total_loss'], rtol=1e-5)
    np.testing.assert_allclose(float(reports['ssim_loss']), 0.0)
    assert_grad_close(host(torch.stack([p.grad for p in poses], 1)), 3.0 * G['gpose'], what='gpose via autograd')
    assert_grad_close(host(disps[2].grad), 3.0 * G['gdisp'][2], what='gdisp via autograd')
    assert_grad_close(host(masks[1].grad), 3.0 * G['glogits'][1], what='glogits via autograd')


@pytest.mark.parametrize('flagset', ['v1_ssim', 'v1_odom'])
def test_back_to_back_calls_are_ordered(flagset):
    """The four kernels of a step are chained with programmatic dependent launches and consecutive calls share
    one workspace: many un-synchronised calls on alternating inputs must reproduce the isolated results bit for
    bit (gdisp / glogits are written without atomics) -- guards the grid-dependency waits."""
    import torch
    flags = FLAGSETS[flagset]
    op = _op(flags)
    sets = [dev_inputs(make_snippets(4, 2, 128, 416, seed=40 + k)) for k in range(2)]
    ref = []
    for g in sets:
        l, gr = op.forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
        torch.cuda.synchronize()
        ref.append((host(l), [host(x) for x in gr['gdisps']], host(gr['gposes'])))
    outs = []
    for it in range(24):
        g = sets[it & 1]
        outs.append(op.forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits']))
    torch.cuda.synchronize()
    for it, (l, gr) in enumerate(outs):
        rl, rg, rp = ref[it & 1]
        np.testing.assert_allclose(host(l), rl, rtol=1e-6)
        for s in range(4):
            np.testing.assert_array_equal(host(gr['gdisps'][s]), rg[s])
        np.testing.assert_allclose(host(gr['gposes']), rp, rtol=1e-5, atol=1e-9)


def test_abi_communicator_world_of_one_inside_a_cuda_graph():
    """sfm_comm_* / sfm_allreduce_partials on one GPU: a world of one leaves the partials unchanged, and the call is
    capturable in a CUDA graph together with the step (the N-GPU equality is tests/test_gpu_multi.py)."""
    import torch
    from sfm_learner_chainer_b200.distributed import LossPartialsComm, ShardedViewSynthesisLoss
    flags = FLAGSETS['v1_odom']
    d = make_snippets(2, 2, 64, 208, seed=91)
    g = dev_inputs(d)
    ref, _ = _op(flags).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    comm = LossPartialsComm(0, 1, exchange=lambda ident: ident)
    op = ShardedViewSynthesisLoss(comm=comm, **flags)
    args = (g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    l1, _, _ = op.forward_backward(*args)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(host(l1), host(ref))
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        op.forward_backward(*args)
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            l2, _, _ = op.forward_backward(*args)
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(host(l2), host(ref))
    comm.close()


def test_peer_loss_sum_world_of_one_inside_a_cuda_graph():
    """sfm_peer_* / sfm_loss_forward_backward_peer on one GPU: the epilogue's in-kernel exchange with a world of one
    returns the local losses, step after step (the device-side step counter) and when replayed as a CUDA graph."""
    import torch
    from sfm_learner_chainer_b200.distributed import PeerLossSum
    flags = FLAGSETS['v1_ssim']
    d = make_snippets(2, 2, 64, 208, seed=92)
    g = dev_inputs(d)
    args = (g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    ref, gref = _op(flags).forward_backward(*args)
    peer = PeerLossSum(0, 1)
    op = _op(flags)
    for _ in range(3):
        l1, g1 = op.forward_backward(*args, peer=peer)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(host(l1), host(ref))
        np.testing.assert_array_equal(host(g1['gposes']), host(gref['gposes']))
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        op.forward_backward(*args, peer=peer)
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            l2, _ = op.forward_backward(*args, peer=peer)
    for _ in range(4):
        graph.replay()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(host(l2), host(ref))
    peer.close()


def test_non_finite_poses_and_huge_depths_are_out_of_view():
    """Documented domain deviation (spec arithmetic): a NaN / Inf pose or a camera-space depth
    beyond ~1e37 makes the pixel OUT OF VIEW (warped value 0, masked, gradient 0) where the reference would propagate
    NaN into the loss.  Pinned here: a (snippet, source) pair with a NaN pose behaves exactly like a pair thrown out of
    view by a huge sideways translation, and a single near-zero (but normal, positive) disparity leaves every output
    finite."""
    flags = FLAGSETS['v1_ssim']
    d = make_snippets(2, 2, 64, 208, seed=93)
    ref = {k: (v.copy() if hasattr(v, 'copy') else [x.copy() for x in v]) for k, v in d.items()}
    ref['poses'][0, 0, :] = 0
    ref['poses'][0, 0, 3] = 1e4                              # every pixel of pair (0, 0) out of view (transform.py:128-131)
    ref['poses'][1, 1, :] = 0
    ref['poses'][1, 1, 3] = 1e4
    bad = {k: (v.copy() if hasattr(v, 'copy') else [x.copy() for x in v]) for k, v in d.items()}
    bad['poses'][0, 0, :] = np.nan
    bad['poses'][1, 1, 4] = np.inf
    out = []
    for dd in (ref, bad):
        g = dev_inputs(dd)
        l, gr = _op(flags).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
        out.append((host(l), [host(x) for x in gr['gdisps']], host(gr['gposes'])))
    assert np.isfinite(out[1][0]).all() and np.isfinite(out[1][2]).all()
    np.testing.assert_array_equal(out[1][0], out[0][0])
    for s in range(4):
        np.testing.assert_array_equal(out[1][1][s], out[0][1][s])
    np.testing.assert_array_equal(out[1][2][0, 1], out[0][2][0, 1])            # the healthy pairs are untouched
    np.testing.assert_array_equal(out[1][2][1, 0], out[0][2][1, 0])
    assert not out[1][2][0, 0].any() and not out[1][2][1, 1].any()              # no gradient through a non-finite pose
    # a disparity of 2e-38 (the smallest normal floats; depth 5e37: the projection overflows) at one pixel
    d['disps'][0][0, 0, 10, 20] = 2e-38
    g = dev_inputs(d)
    op = _op(flags)
    l, gr = op.forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
    assert np.isfinite(host(l)).all() and np.isfinite(host(gr['gposes'])).all()
    assert all(np.isfinite(host(x)).all() for x in gr['gdisps'])
    _, dbg = op.forward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'], debug=True)
    assert not host(dbg['inb'][0])[0, :, 10, 20].any()


def test_ssim_record_placements_agree(monkeypatch):
    """The SSIM kernel keeps its two-row delay line of forward records in registers (sub-wave grids) or in shared memory
    (large grids); the launch policy picks.  Same arithmetic in the same order: per-pixel gradients must be identical
    bit for bit, for the one-warp and the source-split task shapes."""
    flags = FLAGSETS['v1_ssim']
    d = make_snippets(2, 2, 72, 136, seed=94, harsh=True)
    g = dev_inputs(d)
    for nw in ('1', '2'):
        res = []
        for srec in ('0', '1'):
            monkeypatch.setenv('SFM_SSIM_NW', nw)
            monkeypatch.setenv('SFM_SSIM_SREC', srec)
            l, gr = _op(flags).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], g['logits'])
            res.append((host(l), [host(x) for x in gr['gdisps']], host(gr['gposes'])))
        np.testing.assert_allclose(res[1][0], res[0][0], rtol=1e-6)
        assert_grad_close(res[1][2], res[0][2], what='gposes')
        for s in range(4):
            np.testing.assert_array_equal(res[1][1][s], res[0][1][s], err_msg='gdisp scale %d, NW=%s' % (s, nw))


@pytest.mark.parametrize('ns', [1, 2, 3])
def test_fewer_scales_match_oracle(ns):
    """SfmDesc.n_scales < 4 (the ABI allows 1..4): with one scale the prologue kernel has no pyramid to build at all
    (scale 0 is read from the caller's tensors); losses and gradients against the oracle run with the same n_scales."""
    from sfm_learner_chainer_b200 import ViewSynthesisLoss
    flags = FLAGSETS['v1_ssim']
    d = make_snippets(2, 2, 64, 208, seed=95)
    K = np.ascontiguousarray(d['intrinsics'][:, :ns])
    L, G, _ = O.sfm_loss(d['tgt'], d['src'], K, d['disps'][:ns], d['poses'], None, O.LossConfig(n_scales=ns, **flags))
    op = ViewSynthesisLoss(n_scales=ns, **flags)
    losses, grads = op.forward_backward(to_dev(d['tgt']), to_dev(d['src']), to_dev(K), [to_dev(x) for x in d['disps'][:ns]],
                                        to_dev(d['poses']), None)
    np.testing.assert_allclose(host(losses), O.losses_vec(L), rtol=1e-5, atol=1e-9)
    assert_grad_close(host(grads['gposes']), G['gpose'], what='gpose')
    for s in range(ns):
        assert_grad_close(host(grads['gdisps'][s]), G['gdisp'][s], what='gdisp[%d]' % s)
    tp, sp = op.pyramid(to_dev(d['tgt']), to_dev(d['src']))
    assert len(tp) == ns and len(sp) == ns
