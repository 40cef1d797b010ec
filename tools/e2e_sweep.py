"""Development aid: end-to-end (host buffers) throughput for several numbers of host contexts in flight."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
for u8 in (True, False):
    for n in (1, 2, 3, 4, 6, 8, 12, 16):
        r = bench.run_e2e(cfg, dev, n_ctx=n, u8=u8, repeats=2, min_steps=200, min_seconds=0.3)
        print('%s u8=%s n_ctx=%d: %.0f Mpix/s  (%.1f us/step, H2D %.1f GB/s)' % (cfg, u8, n, r['value'],
              bench.CONFIGS[cfg]['B'] * bench.pyramid_pixels(bench.CONFIGS[cfg]['H'], bench.CONFIGS[cfg]['W']) / r['value'], r['h2d_gbs']), flush=True)
