"""Torch stand-ins for the two CNNs either side of the loss path -- TEST / BENCH INFRASTRUCTURE, not product.

The reference's DispNet (models/disp_net.py:16-124) and PoseNet (models/pose_net.py:7-81) are Chainer links and
out of scope for this repository (the loss path is a drop-in underneath them).  Chainer is not installable in
this image, so BASELINE config 3 ("full train step of sfm_learner_v1.yml") is exercised with torch modules of
the same layer shapes (kernel sizes, strides, channel counts, skip connections, multi-scale heads): same tensor
shapes at the seam, same parameter count (~39.9 M at S=2), random init.  Convolutions run in cuDNN.

`raw_seam=True` makes the nets stop one op earlier, the way a drop-in with the producer-side fusion would:
DispNet returns the pre-activation `dispout1` map for scale 0 (disp2..4 stay activated, the decoder consumes
them) and PoseNet returns the raw `poseout` map instead of 0.01 * mean.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

DISP_SCALING, MIN_DISP = 10.0, 0.01


def _resize_like(x, ref):
    if x.shape[2:] == ref.shape[2:]:
        return x
    return F.interpolate(x, size=ref.shape[2:], mode='bilinear', align_corners=True)


class DispNetStandIn(nn.Module):
    def __init__(self, raw_seam=False):
        super().__init__()
        self.raw_seam = raw_seam
        enc = [(3, 32, 7), (32, 64, 5), (64, 128, 3), (128, 256, 3), (256, 512, 3), (512, 512, 3), (512, 512, 3)]
        self.down = nn.ModuleList([nn.Conv2d(i, o, k, 2, k // 2) for i, o, k in enc])
        self.same = nn.ModuleList([nn.Conv2d(o, o, k, 1, k // 2) for _, o, k in enc])
        dec = [(512, 512, 512), (512, 512, 512), (512, 256, 256), (256, 128, 128), (128, 64, 64 + 1), (64, 32, 32 + 1), (32, 16, 1)]
        self.up = nn.ModuleList([nn.ConvTranspose2d(i, o, 4, 2, 1) for i, o, _ in dec])
        self.merge = nn.ModuleList([nn.Conv2d(o + skip, o, 3, 1, 1) for _, o, skip in dec])
        self.heads = nn.ModuleList([nn.Conv2d(c, 1, 3, 1, 1) for c in (128, 64, 32, 16)])   # dispout4 .. dispout1

    def forward(self, x):
        H, W = x.shape[2:]
        skips, h = [], x
        for d, s in zip(self.down, self.same):
            h = F.relu(s(F.relu(d(h))))
            skips.append(h)
        outs, raws, prev_up = [], [], None
        for k in range(7):
            h = F.relu(self.up[k](h))
            if k < 6:
                skip = skips[5 - k]
                h = _resize_like(h, skip)
                cat = [h, skip] if prev_up is None else [h, skip, prev_up]
            else:
                cat = [h, prev_up]
            h = F.relu(self.merge[k](torch.cat(cat, 1)))
            prev_up = None
            if k >= 3:
                raw = self.heads[k - 3](h)
                disp = DISP_SCALING * torch.sigmoid(raw) + MIN_DISP
                outs.append(disp)
                raws.append(raw)
                if k < 6:
                    size = (H // 4, W // 4) if k == 3 else (H // 2, W // 2) if k == 4 else (H, W)
                    prev_up = F.interpolate(disp, size=size, mode='bilinear', align_corners=True)
        outs, raws = outs[::-1], raws[::-1]            # [disp1 (full res), disp2, disp3, disp4]
        if self.raw_seam:
            outs[0] = raws[0]
        return outs


class PoseNetStandIn(nn.Module):
    def __init__(self, n_sources=2, raw_seam=False):
        super().__init__()
        self.n_sources, self.raw_seam = n_sources, raw_seam
        cin = 3 * (1 + n_sources)
        spec = [(cin, 16, 7), (16, 32, 5), (32, 64, 3), (64, 128, 3), (128, 256, 3)]
        self.enc = nn.ModuleList([nn.Conv2d(i, o, k, 2, k // 2) for i, o, k in spec])
        self.pose1 = nn.Conv2d(256, 256, 3, 2, 1)
        self.pose2 = nn.Conv2d(256, 256, 3, 2, 1)
        self.poseout = nn.Conv2d(256, 6 * n_sources, 1)
        self.exp = nn.ModuleList([nn.ConvTranspose2d(256, 256, 4, 2, 1), nn.ConvTranspose2d(256, 128, 4, 2, 1),
                                  nn.ConvTranspose2d(128, 64, 4, 2, 1), nn.ConvTranspose2d(64, 32, 6, 2, 2),
                                  nn.ConvTranspose2d(32, 16, 6, 2, 2)])
        self.expout = nn.ModuleList([nn.Conv2d(128, n_sources, 3, 1, 1), nn.Conv2d(64, n_sources, 3, 1, 1),
                                     nn.Conv2d(32, n_sources, 5, 1, 2), nn.Conv2d(16, n_sources, 7, 1, 3)])

    def forward(self, tgt, stacked_src, do_exp=True):
        h = torch.cat([tgt, stacked_src], 1)
        for c in self.enc:
            h = F.relu(c(h))
        p = self.poseout(F.relu(self.pose2(F.relu(self.pose1(h)))))
        if self.raw_seam:
            poses = p                                                    # (B, 6S, h', w')
        else:
            m = 0.01 * p.mean((2, 3))
            poses = tuple(torch.split(m, 6, dim=1))
        masks = None
        if do_exp:
            e = F.relu(self.exp[0](h))
            masks = []
            for k in range(1, 5):
                e = F.relu(self.exp[k](e))
                masks.append(self.expout[k - 1](e))
            masks = masks[::-1]
        return poses, masks
