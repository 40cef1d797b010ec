"""Development aid: times the step (CUDA-graph replay) and the fused kernel of BASELINE shapes under a list of
development knobs (environment variables read by the launchers at capture time), one Workload per config.
usage: python tools/exp_variants.py cfg4 cfg5 -- "" "SFM_LIFO=1" "SFM_PF=2960" "SFM_LIFO=1 SFM_PF=2960"
"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

args = sys.argv[1:]
cut = args.index('--')
cfgs, variants = args[:cut], args[cut + 1:]
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
peak, _ = bench.measured_peak()
KNOBS = set()
for v in variants:
    for kv in v.split():
        KNOBS.add(kv.split('=')[0])
for name in cfgs:
    blocal = None
    if ':' in name:                      # cfg5:8 = the cfg5 shape with 8 snippets on this GPU (a shard of the global batch)
        name, blocal = name.split(':')[0], int(name.split(':')[1])
    wl = bench.Workload(name, dev, B_global=bench.CONFIGS[name]['B'] if blocal else None, B_local=blocal)
    n = 100 if name != 'cfg5' else 20
    for v in variants:
        for k in KNOBS:
            os.environ.pop(k, None)
        for kv in v.split():
            k, val = kv.split('=')
            os.environ[k] = val
        wl.capture()
        ms, _, _ = wl.time_steps(n, 5)
        ms2, _, _ = wl.time_steps(n, 2)
        km, kmed = wl.time_fused_kernel(n)
        print(json.dumps(dict(cfg=name, knobs=v, step_us=round(min(ms, ms2) * 1e3, 2), fused_us=round(kmed * 1e3, 2),
                              step_frac=round(wl.A_strict / min(ms, ms2) / 1e6 / peak, 4))), flush=True)
    for k in KNOBS:
        os.environ.pop(k, None)
    del wl
    torch.cuda.empty_cache()
