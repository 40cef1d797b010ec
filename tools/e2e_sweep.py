"""End-to-end (host-buffer) throughput of the loss path at cfg2 against the number of host contexts kept in flight,
for float images and for uint8 frames (bench.py's run_e2e)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
import torch, bench
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
for n in (1, 2, 3, 4):
    v, h2d, d2h, _ = bench.run_e2e('cfg2', 300, dev, n_ctx=n)
    print('n_ctx', n, round(v, 1), 'Mpix/s', round(282880 / v, 1), 'us/step')
for n in (2, 3):
    v, h2d, d2h, _ = bench.run_e2e('cfg2', 300, dev, n_ctx=n, u8=True)
    print('u8 n_ctx', n, round(v, 1), 'Mpix/s', round(282880 / v, 1), 'us/step', h2d)
