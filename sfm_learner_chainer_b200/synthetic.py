"""Seeded synthetic KITTI-shaped snippets for tests and bench (numpy only).

Shapes and value ranges follow the reference's data layer:
images are ``img/127.5 - 1`` in [-1, 1] (datasets/kitti/kitti_raw_dataset.py:12-14),
intrinsics are a 4-level pyramid ``K / 2^s`` (datasets/kitti/kitti_raw_transformed.py:76-93),
disparities are ``10*sigmoid(x) + 0.01`` (models/disp_net.py:7-8,104) and poses
are ``0.01 * mean(...)`` sized (models/pose_net.py:52).
"""
import numpy as np

N_SCALES = 4


def _upsample(lo, H, W):
    """align-corners bilinear upsample of (..., h, w) to (..., H, W)."""
    h, w = lo.shape[-2:]
    ys = np.linspace(0, h - 1, H)
    xs = np.linspace(0, w - 1, W)
    y0 = np.clip(np.floor(ys).astype(int), 0, h - 2)
    x0 = np.clip(np.floor(xs).astype(int), 0, w - 2)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    a = lo[..., y0[:, None], x0[None, :]]
    b = lo[..., y0[:, None], x0[None, :] + 1]
    c = lo[..., y0[:, None] + 1, x0[None, :]]
    d = lo[..., y0[:, None] + 1, x0[None, :] + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def _smooth_field(rs, shape, H, W, cell=8):
    h, w = max(H // cell, 2), max(W // cell, 2)
    return _upsample(rs.standard_normal(shape + (h, w)), H, W)


def make_intrinsics(B, H, W, n_scales=N_SCALES, dtype=np.float32):
    """KITTI P_rect scaled to (H, W) and its K/2^s pyramid -> (B, n_scales, 3, 3)."""
    fx, fy = 241.67 * W / 416.0, 246.28 * H / 128.0
    cx, cy = 204.2 * W / 416.0, 59.0 * H / 128.0
    K = np.zeros((B, n_scales, 3, 3), dtype)
    for s in range(n_scales):
        K[:, s] = np.array([[fx / 2 ** s, 0, cx / 2 ** s],
                            [0, fy / 2 ** s, cy / 2 ** s],
                            [0, 0, 1]], dtype)
    return K


def make_snippets(B, S, H, W, seed=0, n_scales=N_SCALES, harsh=False,
                  rough_disp=False, dtype=np.float32):
    """Returns dict(tgt, src, intrinsics, disps, poses, logits).

    tgt (B,3,H,W), src (B,S,3,H,W), intrinsics (B,n_scales,3,3),
    disps list of (B,1,h_s,w_s), poses (B,S,6), logits list of (B,S,h_s,w_s).
    harsh=True widens the pose range (about 30% of pixels leave the view);
    rough_disp=True makes the disparity i.i.d. per pixel instead of smooth.
    """
    rs = np.random.RandomState(seed)

    def image(shape):
        h, w = max(H // 8, 2), max(W // 8, 2)
        lo = rs.uniform(-1, 1, shape + (h, w))
        img = 0.9 * _upsample(lo, H, W) + rs.uniform(-0.1, 0.1, shape + (H, W))
        img = np.where(np.abs(img) < 1e-6, 1e-3, img)       # no exact zeros (base_model.py:96)
        return np.ascontiguousarray(img, dtype=dtype)      # fancy indexing above leaves a permuted memory order

    tgt = image((B, 3))
    src = image((B, S, 3))
    disps, logits = [], []
    for s in range(n_scales):
        h, w = H >> s, W >> s
        if rough_disp:
            x = rs.standard_normal((B, 1, h, w))
        else:
            x = _smooth_field(rs, (B, 1), h, w) + 0.05 * rs.standard_normal((B, 1, h, w))
        disps.append(np.ascontiguousarray(10.0 / (1.0 + np.exp(-x)) + 0.01, dtype=dtype))
        logits.append(np.ascontiguousarray(rs.standard_normal((B, S, h, w)), dtype=dtype))
    if harsh:
        r = rs.uniform(-0.02, 0.02, (B, S, 3))
        t = rs.uniform(-0.05, 0.05, (B, S, 3))
    else:
        r = rs.uniform(-0.01, 0.01, (B, S, 3))
        t = rs.uniform(-0.01, 0.01, (B, S, 3))
    poses = np.ascontiguousarray(np.concatenate([r, t], axis=-1), dtype=dtype)
    return dict(tgt=tgt, src=src, intrinsics=make_intrinsics(B, H, W, n_scales, dtype),
                disps=disps, poses=poses, logits=logits)


def make_raw_seam(data, pose_hw=(1, 4), seed=0):
    """Pre-activation inputs of the seam either side of the loss for the snippets `data` of make_snippets:
    raw disparity maps x with 10*sigmoid(x)+0.01 ~ data['disps'] (models/disp_net.py:104) and a raw `poseout`
    map (B, 6*S, h', w') whose 0.01*mean over (h', w') ~ data['poses'] (models/pose_net.py:52)."""
    rs = np.random.RandomState(seed + 1000)
    dtype = data['tgt'].dtype
    raw_disps = []
    for d in data['disps']:
        y = np.clip((d.astype(np.float64) - 0.01) / 10.0, 1e-6, 1 - 1e-6)
        raw_disps.append(np.ascontiguousarray(np.log(y / (1 - y)), dtype=dtype))
    B, S = data['poses'].shape[:2]
    ph, pw = pose_hw
    centre = data['poses'].astype(np.float64).reshape(B, 6 * S, 1, 1) / 0.01
    noise = rs.standard_normal((B, 6 * S, ph, pw))
    noise -= noise.mean(axis=(2, 3), keepdims=True)
    raw_pose = np.ascontiguousarray(centre + 0.5 * np.abs(centre).mean() * noise, dtype=dtype)
    return raw_disps, raw_pose


# BASELINE.json configs -> concrete shapes and reference flag sets (experiments/*.yml)
CONFIGS = {
    'cfg1': dict(B=4, S=2, H=128, W=416, smooth_reg=0.0, exp_reg=0.0, ssim_rate=0.0),    # sfm_learner_v1.yml:14-16
    'cfg2': dict(B=4, S=2, H=128, W=416, smooth_reg=0.1, exp_reg=0.0, ssim_rate=0.15),   # sfm_learner_v1_ssim.yml:14-17
    'cfg4': dict(B=32, S=4, H=128, W=416, smooth_reg=0.1, exp_reg=0.2, ssim_rate=0.0),   # sfm_learner_v1_odom.yml:14-16
    'cfg5': dict(B=64, S=2, H=256, W=832, smooth_reg=0.1, exp_reg=0.0, ssim_rate=0.15),  # high-res stress, cfg2 flags
}
