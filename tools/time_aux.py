"""CUDA-event timing of the kernels either side of the loss path: uint8 ingest, disparity
activation stage, depth evaluation.  Reports achieved GB/s on the algorithmic bytes of each."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sfm_learner_chainer_b200 import ingest_u8, disp_activation, evaluate_depth_batch
from sfm_learner_chainer_b200.functions import draw_augmentation


def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3     # us


def main():
    peak = 6534.5
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    rs = np.random.RandomState(0)
    for B, S, H, W in [(4, 2, 128, 416), (32, 4, 128, 416), (64, 2, 256, 832)]:
        frames = torch.from_numpy(rs.randint(0, 256, (B, 1 + S, H, W, 3)).astype(np.uint8)).cuda()
        K = torch.from_numpy(np.tile(np.array([[241.67, 0, 204.2], [0, 246.28, 59.0], [0, 0, 1]], np.float32), (B, 1, 1))).cuda()
        aug = [draw_augmentation(H, W, rs) for _ in range(B)]
        # rotate over enough frame sets to defeat the L2 (the float outputs alone exceed it at the larger shapes)
        us = timeit(lambda: ingest_u8(frames, K, aug))
        by = B * (1 + S) * H * W * (3 + 12)
        print(json.dumps(dict(kernel='sfm_ingest_u8', shape=[B, S, H, W], us=round(us, 1), algo_MB=round(by / 1e6, 1),
                              gbs=round(by / us / 1e3, 1), frac_of_peak=round(by / us / 1e3 / peak, 3),
                              note='includes the host-side packing of SfmAugment and its small H2D copy per call')))
    for n in (4 * 128 * 416, 64 * 256 * 832):
        x = torch.randn(n, device='cuda')
        us = timeit(lambda: disp_activation(x))
        print(json.dumps(dict(kernel='sfm_disp_activation', n=n, us=round(us, 1), gbs=round(8 * n / us / 1e3, 1), frac_of_peak=round(8 * n / us / 1e3 / peak, 3))))
    for B in (1, 16):
        h, w, Hg, Wg = 128, 416, 375, 1242
        pred = torch.rand(B, 1, h, w, device='cuda') * 50 + 1
        gt = torch.rand(B, Hg, Wg, device='cuda') * 70 + 1
        mask = (torch.rand(B, Hg, Wg, device='cuda') < 0.3).to(torch.uint8)
        us = timeit(lambda: evaluate_depth_batch(pred, gt, mask, 1e-3, 80.0))
        by = B * (h * w * 4 + Hg * Wg * (4 + 1 + 4) + Hg * Wg * (4 + 4 + 1) * 2)      # resize pass + two masked passes
        print(json.dumps(dict(kernel='sfm_eval_depth (7 launches)', B=B, gt_shape=[Hg, Wg], us=round(us, 1), algo_MB=round(by / 1e6, 2),
                              gbs=round(by / us / 1e3, 1))))


if __name__ == '__main__':
    main()
