import sys, os, json
sys.path.insert(0, '.')
import torch, bench
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
for name in sys.argv[1:]:
    wl = bench.Workload(name, dev)
    wl.capture()
    ms, _, _ = wl.time_steps(300 if name != 'cfg5' else 30, 10)
    pms = wl.time_pipelined(300 if name != 'cfg5' else 30, 10)
    print(json.dumps(dict(cfg=name, step_us=round(ms*1e3, 2), pipelined_us=round(pms*1e3, 2))))
    del wl; torch.cuda.empty_cache()
