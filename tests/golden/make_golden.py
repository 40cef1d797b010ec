#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the REFERENCE'S OWN source files.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It puts `oracle/chainer_shim` (torch-CPU stand-in for the un-installable
chainer==4.0.0b1, see oracle/chainer_shim/README.md) and `/root/reference` on
sys.path, imports the unmodified `models.base_model`, `models.transform` and
`models.spational_transformer_sampler_interp`, replaces only the two CNNs
(out of scope; `disp_net` / `pose_net` attributes) with stubs that return the
seeded synthetic predictions, and records inputs, the five reported losses and
the gradients w.r.t. disparities, poses and explainability logits.

The fixtures travel with the repo; nothing at test time reads /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('SFM_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'chainer_shim'))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import chainer                                                    # the shim
from chainer import Variable
import models.transform as ref_transform                          # reference code, unmodified
import models.base_model as ref_base_model
import models.spational_transformer_sampler_interp as ref_interp
from sfm_learner_chainer_b200.synthetic import make_snippets

assert chainer.__version__.endswith('shim')

CASES = [
    # name, B, S, H, W, seed, harsh, flags (experiments/sfm_learner_v1*.yml architecture blocks)
    ('v1',        2, 2, 32, 104, 0, False, dict(smooth_reg=0.0, exp_reg=0.0, seq_len=3)),
    ('v1_ssim',   2, 2, 32, 104, 1, False, dict(smooth_reg=0.1, exp_reg=0, seq_len=3, ssim_rate=0.15)),
    ('v1_odom',   2, 4, 32, 104, 2, False, dict(smooth_reg=0.1, exp_reg=0.2, seq_len=5)),
    ('v1_ssim_harsh', 1, 2, 48, 160, 3, True, dict(smooth_reg=0.1, exp_reg=0, seq_len=3, ssim_rate=0.15)),
    ('v1_odom_harsh', 1, 4, 48, 160, 4, True, dict(smooth_reg=0.1, exp_reg=0.2, seq_len=5)),
]


def reset_reference_caches():
    ref_transform.filler = None        # transform.py:62 (keyed on N only)
    ref_transform.meshgrid = None      # transform.py:135 (keyed on H*W only)


def run_model(data, flags, dtype):
    reset_reference_caches()
    chainer.clear_reports()
    model = ref_base_model.SFMLearner(flags, {'download': None, 'path': None})
    cast = lambda a: np.ascontiguousarray(a.astype(dtype))
    disps = [Variable(cast(d)) for d in data['disps']]
    S = data['src'].shape[1]
    poses = [Variable(cast(data['poses'][:, i])) for i in range(S)]
    masks = [Variable(cast(l)) for l in data['logits']]
    model.disp_net = lambda tgt: disps
    model.pose_net = lambda tgt, src, do_exp=True: (tuple(poses), masks if do_exp else None)
    K = cast(data['intrinsics'])
    loss = model(cast(data['tgt']), cast(data['src']), K, K)
    loss.backward()
    rep = chainer.get_reports()
    out = {}
    for k in ('total_loss', 'pixel_loss', 'smooth_loss', 'exp_loss', 'ssim_loss'):
        v = rep[k]
        out[k] = float(v.data) if isinstance(v, Variable) else float(v)
    out['gdisp'] = [d.grad for d in disps]
    out['gpose'] = np.stack([p.grad for p in poses], axis=1)
    do_exp = flags['exp_reg'] is not None and flags['exp_reg'] > 0
    out['glogits'] = [m.grad for m in masks] if do_exp else None
    return out


def run_model_raw(data, raw_disps, raw_pose, flags, dtype):
    """Same as run_model with the seam on either side of the loss included: the disparities are formed from
    pre-activation maps with DispNet's expression (models/disp_net.py:104, the reference's own DISP_SCALING /
    MIN_DISP constants) and the poses by running the reference's own PoseNet.pred_pose (pose_net.py:47-54)
    on a stand-in `self` whose three convolutions are the identity."""
    import models.disp_net as ref_disp_net
    import models.pose_net as ref_pose_net
    F = chainer.functions
    reset_reference_caches()
    chainer.clear_reports()
    model = ref_base_model.SFMLearner(flags, {'download': None, 'path': None})
    cast = lambda a: np.ascontiguousarray(a.astype(dtype))
    xs = [Variable(cast(d)) for d in raw_disps]
    xp = Variable(cast(raw_pose))
    S = data['src'].shape[1]
    masks = [Variable(cast(l)) for l in data['logits']]

    class FakePoseNet(object):
        n_sources = S
        activation = staticmethod(lambda h: h)
        pose1 = pose2 = poseout = staticmethod(lambda h: h)

    def disp_net(tgt):
        return [ref_disp_net.DISP_SCALING * F.sigmoid(x) + ref_disp_net.MIN_DISP for x in xs]

    def pose_net(tgt, src, do_exp=True):
        poses = ref_pose_net.PoseNet.pred_pose(FakePoseNet(), xp)
        return tuple(poses), masks if do_exp else None

    model.disp_net, model.pose_net = disp_net, pose_net
    K = cast(data['intrinsics'])
    loss = model(cast(data['tgt']), cast(data['src']), K, K)
    loss.backward()
    rep = chainer.get_reports()
    out = {}
    for k in ('total_loss', 'pixel_loss', 'smooth_loss', 'exp_loss', 'ssim_loss'):
        v = rep[k]
        out[k] = float(v.data) if isinstance(v, Variable) else float(v)
    out['gx'] = [x.grad for x in xs]
    out['gxpose'] = xp.grad
    do_exp = flags['exp_reg'] is not None and flags['exp_reg'] > 0
    out['glogits'] = [m.grad for m in masks] if do_exp else None
    return out


SEAM_CASES = [
    ('seam_ssim', 2, 2, 32, 104, 5, (1, 4), dict(smooth_reg=0.1, exp_reg=0, seq_len=3, ssim_rate=0.15)),
    ('seam_odom', 1, 4, 32, 104, 6, (2, 5), dict(smooth_reg=0.1, exp_reg=0.2, seq_len=5)),
]


def make_seam():
    from sfm_learner_chainer_b200.synthetic import make_raw_seam
    for name, B, S, H, W, seed, pose_hw, flags in SEAM_CASES:
        data = make_snippets(B, S, H, W, seed=seed)
        raw_disps, raw_pose = make_raw_seam(data, pose_hw, seed=seed)
        blob = dict(tgt=data['tgt'], src=data['src'], intrinsics=data['intrinsics'], raw_pose=raw_pose,
                    flags=np.array([flags['smooth_reg'], flags['exp_reg'] or 0.0, flags.get('ssim_rate', 0.0)]))
        for s in range(4):
            blob['raw_disp%d' % s] = raw_disps[s]
            blob['logits%d' % s] = data['logits'][s]
        for tag, dtype in (('f64', np.float64), ('f32', np.float32)):
            out = run_model_raw(data, raw_disps, raw_pose, flags, dtype)
            blob['losses_' + tag] = np.array([out[k] for k in ('total_loss', 'pixel_loss', 'smooth_loss',
                                                               'exp_loss', 'ssim_loss')], np.float64)
            blob['gxpose_' + tag] = out['gxpose']
            for s in range(4):
                blob['gx%d_%s' % (s, tag)] = out['gx'][s]
                if out['glogits'] is not None:
                    blob['glogits%d_%s' % (s, tag)] = out['glogits'][s]
            print(name, tag, blob['losses_' + tag])
        np.savez_compressed(os.path.join(HERE, '%s.npz' % name), **blob)


def make_edge_smooth():
    """compute_disp_smooth (base_model.py:144-155), the edge-aware smoothness the reference keeps commented out at
    its call site (:78-80): the reference's own method on the pyramid levels, weighted as that comment writes it,
    value and gradient w.r.t. the disparities (float64 and float32)."""
    F = chainer.functions
    B, S, H, W = 2, 2, 32, 104
    data = make_snippets(B, S, H, W, seed=8, rough_disp=True)
    flags = dict(smooth_reg=0.1, exp_reg=0.0, seq_len=3)
    model = ref_base_model.SFMLearner(flags, {'download': None, 'path': None})
    blob = dict(tgt=data['tgt'], smooth_reg=np.float64(flags['smooth_reg']))
    for tag, dtype in (('f64', np.float64), ('f32', np.float32)):
        disps = [Variable(np.ascontiguousarray(d.astype(dtype))) for d in data['disps']]
        loss = 0
        for ns in range(4):
            cur = F.resize_images(data['tgt'].astype(dtype), (H // 2 ** ns, W // 2 ** ns)).data
            loss += (flags['smooth_reg'] / (2 ** ns)) * model.compute_disp_smooth(cur, disps[ns])
        loss.backward()
        blob['loss_' + tag] = np.float64(loss.data)
        for ns in range(4):
            blob['disp%d' % ns] = data['disps'][ns]
            blob['gdisp%d_%s' % (ns, tag)] = disps[ns].grad
        print('edge_smooth', tag, float(loss.data))
    np.savez_compressed(os.path.join(HERE, 'edge_smooth.npz'), **blob)


def make_eval():
    """evaluate.py:94-103 restated line by line with the shim's F.resize_images / F.clip, followed by the reference's
    own kitti_eval.depth_util.compute_depth_errors (pure numpy, imported unmodified)."""
    import kitti_eval.depth_util as ref_du
    F = chainer.functions
    rs = np.random.RandomState(9)
    B, h, w, Hg, Wg = 2, 32, 104, 94, 311
    min_depth, max_depth = 1e-3, 80.0
    gt = (rs.uniform(1.0, 60.0, (B, Hg // 8 + 2, Wg // 8 + 2)).repeat(8, 1).repeat(8, 2)[:, :Hg, :Wg]
          * rs.uniform(0.9, 1.1, (B, Hg, Wg))).astype(np.float32)
    pred_depth = (0.37 * gt[:, ::3, ::3][:, :h, :w] * rs.uniform(0.7, 1.4, (B, h, w))).astype(np.float32)[:, None]
    pred_depth[0, 0, :2] = 1e-5            # below min_depth: clipped
    pred_depth[1, 0, -2:] = 500.0          # above max_depth after scaling irrelevant: clipped before
    mask = (rs.uniform(0, 1, (B, Hg, Wg)) < 0.4)
    mask[:, :Hg // 3] = False              # KITTI crop: no ground truth in the sky
    out = []
    for even in (True, False):
        mk = mask.copy()
        if (int(mk.sum()) % 2 == 0) != even:
            mk[tuple(np.argwhere(mk)[0])] = False
        pd = F.resize_images(pred_depth, gt.shape[1:]).data                  # evaluate.py:94
        pd = F.clip(pd, min_depth, max_depth).data[:, 0]                      # :95
        pdm, gtm = pd[mk], gt[mk]                                             # :99-100
        scale_factor = np.median(gtm) / np.median(pdm)                        # :101
        pdm = pdm * scale_factor                                              # :102 (pred_depth *= scale_factor)
        out.append((mk, ref_du.compute_depth_errors(gtm, pdm), scale_factor))
        print('eval', 'even' if even else 'odd', int(mk.sum()), out[-1][1], scale_factor)
    np.savez_compressed(os.path.join(HERE, 'eval_depth.npz'), pred_depth=pred_depth, gt=gt, min_depth=min_depth,
                        max_depth=max_depth, mask_even=out[0][0], errors_even=out[0][1], scale_even=out[0][2],
                        mask_odd=out[1][0], errors_odd=out[1][1], scale_odd=out[1][2])


def make_ingest():
    """The data layer in front of the loss: the reference's own load_as_float_norm
    (datasets/kitti/kitti_raw_dataset.py:12-14) and _transform = data_augmentation + get_multi_scale_intrinsics
    (datasets/kitti/kitti_raw_transformed.py:23-102) on synthetic uint8 frames, with the global numpy RNG seeded
    per snippet.  The modules' unrelated imports (cv2, PIL, scipy.misc.imread, chainer.dataset(s)) are stubbed;
    imread is the stub that hands the synthetic frame to load_as_float_norm."""
    import types
    frames_by_path = {}
    for name in ('cv2', 'PIL', 'PIL.Image'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['PIL'].Image = sys.modules['PIL.Image']
    import scipy
    misc = types.ModuleType('scipy.misc')
    misc.imread = lambda path: frames_by_path[path]
    sys.modules['scipy.misc'] = misc
    scipy.misc = misc
    ds = types.ModuleType('chainer.datasets')
    ds.TransformDataset = type('TransformDataset', (object,), {'__init__': lambda self, d, t: None})
    dsm = types.ModuleType('chainer.dataset')
    dsm.DatasetMixin = type('DatasetMixin', (object,), {})
    sys.modules['chainer.datasets'], sys.modules['chainer.dataset'] = ds, dsm
    chainer.datasets, chainer.dataset = ds, dsm
    import datasets.kitti.kitti_raw_dataset as ref_ds                # reference code, unmodified
    import datasets.kitti.kitti_raw_transformed as ref_tr

    B, S, H, W = 3, 2, 32, 104
    rs = np.random.RandomState(77)
    lo = rs.uniform(0, 255, (B, 1 + S, H // 4, W // 4, 3))
    frames = np.clip(np.kron(lo, np.ones((1, 1, 4, 4, 1))) + rs.uniform(-20, 20, (B, 1 + S, H, W, 3)), 0, 255).astype(np.uint8)
    K = np.array([[[241.67 * W / 416, 0, 204.2 * W / 416], [0, 246.28 * H / 128, 59.0 * H / 128], [0, 0, 1]]] * B, np.float32)
    K[1] *= np.array([[1.03, 1, 0.98], [1, 0.97, 1.05], [1, 1, 1]], np.float32)
    seeds = np.array([11, 12, 14])               # 14 flips, 11 and 12 do not (checked below)
    tgt, src, Ks, flips = [], [], [], []
    for b in range(B):
        for j in range(1 + S):
            frames_by_path['f%d_%d' % (b, j)] = frames[b, j]
        t = ref_ds.load_as_float_norm('f%d_0' % b)
        r = [ref_ds.load_as_float_norm('f%d_%d' % (b, j)) for j in range(1, 1 + S)]
        np.random.seed(int(seeds[b]))
        to, so, Ko, _ = ref_tr._transform((t, r, np.copy(K[b]), np.linalg.inv(K[b])), n_scale=4)
        tgt.append(np.asarray(to))
        src.append(np.asarray(so))
        Ks.append(np.stack(Ko))
    # un-augmented path (validation split): load_as_float_norm + get_multi_scale_intrinsics only
    plain_tgt = np.stack([ref_ds.load_as_float_norm('f%d_0' % b) for b in range(B)])
    plain_K = np.stack([np.stack(ref_tr.get_multi_scale_intrinsics(K[b], 4)) for b in range(B)])
    np.savez_compressed(os.path.join(HERE, 'ingest_u8.npz'), frames=frames, K=K, seeds=seeds, tgt=np.stack(tgt),
                        src=np.stack(src), intrinsics=np.stack(Ks), plain_tgt=plain_tgt, plain_intrinsics=plain_K)
    print('ingest_u8', np.stack(tgt).shape, np.stack(src).shape, np.stack(Ks).shape, np.stack(tgt).dtype)


def run_warp(data, dtype, scale=0, i=0):
    """projective_inverse_warp (transform.py:156) + its stages at one scale."""
    reset_reference_caches()
    cast = lambda a: np.ascontiguousarray(a.astype(dtype))
    B, S, _, H, W = data['src'].shape
    h, w = H >> scale, W >> scale
    stacked = cast(data['src']).reshape(B, 3 * S, H, W)
    img = chainer.functions.resize_images(stacked, (h, w)).data[:, 3 * i:3 * i + 3]
    depth = 1.0 / cast(data['disps'][scale])
    depth3 = np.broadcast_to(depth.reshape(B, 1, -1), (B, 3, h * w))
    K = cast(data['intrinsics'][:, scale])
    pose = cast(data['poses'][:, i])
    P = ref_transform.projective_inverse_warp(img, Variable(np.ascontiguousarray(depth3)), Variable(pose), K)
    reset_reference_caches()
    proj = ref_transform.proj_tgt_to_src(Variable(pose), K, B)
    pix = ref_transform.generate_2dmeshgrid(h, w, B)
    cam = ref_transform.pixel2cam(Variable(np.ascontiguousarray(depth3)), pix, K, img.shape)
    grid = ref_transform.cam2pixel(cam, proj, img.shape)
    return dict(img=img, P=P.data, proj=proj.data, grid=grid.data)


def run_interp(seed):
    """SpatialTransformerSamplerInterp forward/backward (pixel-unit grid, clamped)."""
    rs = np.random.RandomState(seed)
    B, C, H, W, oh, ow = 2, 3, 12, 20, 9, 17
    x = rs.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    grid = np.stack([rs.uniform(-3, W + 2, (B, oh, ow)), rs.uniform(-3, H + 2, (B, oh, ow))], 1).astype(np.float32)
    gy = rs.uniform(-1, 1, (B, C, oh, ow)).astype(np.float32)
    xv, gv = Variable(x), Variable(grid)
    y = ref_interp.spatial_transformer_sampler_interp(xv, gv)
    y._t.backward(__import__('torch').from_numpy(gy))
    return dict(x=x, grid=grid, gy=gy, y=y.data, gx=xv.grad, ggrid=gv.grad)


def main():
    if '--seam-only' in sys.argv:
        make_seam()
        return
    if '--edge-only' in sys.argv:
        make_edge_smooth()
        return
    if '--eval-only' in sys.argv:
        make_eval()
        return
    if '--ingest-only' in sys.argv:
        make_ingest()
        return
    for name, B, S, H, W, seed, harsh, flags in CASES:
        data = make_snippets(B, S, H, W, seed=seed, harsh=harsh, rough_disp=(seed % 2 == 1))
        blob = dict(tgt=data['tgt'], src=data['src'], intrinsics=data['intrinsics'], poses=data['poses'],
                    flags=np.array([flags['smooth_reg'], flags['exp_reg'] or 0.0, flags.get('ssim_rate', 0.0)]))
        for s in range(4):
            blob['disp%d' % s] = data['disps'][s]
            blob['logits%d' % s] = data['logits'][s]
        for tag, dtype in (('f64', np.float64), ('f32', np.float32)):
            out = run_model(data, flags, dtype)
            blob['losses_' + tag] = np.array([out[k] for k in ('total_loss', 'pixel_loss', 'smooth_loss',
                                                               'exp_loss', 'ssim_loss')], np.float64)
            blob['gpose_' + tag] = out['gpose']
            for s in range(4):
                blob['gdisp%d_%s' % (s, tag)] = out['gdisp'][s]
                if out['glogits'] is not None:
                    blob['glogits%d_%s' % (s, tag)] = out['glogits'][s]
            print(name, tag, blob['losses_' + tag])
        for sc in (0, 2):
            wout = run_warp(data, np.float64, scale=sc, i=S - 1)
            blob['warp_s%d_P_f64' % sc] = wout['P']
            blob['warp_s%d_grid_f64' % sc] = wout['grid']
            blob['warp_s%d_proj_f64' % sc] = wout['proj']
            blob['warp_s%d_img_f64' % sc] = wout['img']
        np.savez_compressed(os.path.join(HERE, 'loss_%s.npz' % name), **blob)
    np.savez_compressed(os.path.join(HERE, 'interp_sampler.npz'), **run_interp(7))
    make_seam()
    make_ingest()
    make_edge_smooth()
    make_eval()
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
