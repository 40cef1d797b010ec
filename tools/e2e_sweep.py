import sys, os
sys.path.insert(0, '/root/repo')
os.chdir('/root/repo')
import torch, bench
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
for n in (1, 2, 3, 4):
    v, h2d, d2h, _ = bench.run_e2e('cfg2', 300, dev, n_ctx=n)
    print('n_ctx', n, round(v, 1), 'Mpix/s', round(282880 / v, 1), 'us/step')
for n in (2, 3):
    v, h2d, d2h, _ = bench.run_e2e('cfg2', 300, dev, n_ctx=n, u8=True)
    print('u8 n_ctx', n, round(v, 1), 'Mpix/s', round(282880 / v, 1), 'us/step', h2d)
