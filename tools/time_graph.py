"""Device timing via bench.Workload (CUDA-graph replay + events around the fused kernel)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
peak, _ = bench.measured_peak()
for name in (sys.argv[1:] or ['cfg1', 'cfg2', 'cfg4', 'cfg5']):
    wl = bench.Workload(name, dev)
    wl.capture()
    ms, _, _ = wl.time_steps(100 if name != 'cfg5' else 20, 5)
    km, kmed = wl.time_fused_kernel(100 if name != 'cfg5' else 20)
    print(json.dumps(dict(cfg=name, step_us=round(ms * 1e3, 2), fused_us=round(km * 1e3, 2), mpix_s=round(wl.pix / ms / 1e3, 1),
                          step_frac=round(wl.A_strict / ms / 1e6 / peak, 4), fused_frac=round(wl.A_kernel / km / 1e6 / peak, 4))))
    del wl
    torch.cuda.empty_cache()
