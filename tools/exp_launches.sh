#!/bin/bash
# per-kernel durations (ncu, serialised, cold) of the direct-call step for the given configs
mkdir -p gpurun_out
for cfg in "$@"; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launch_$cfg.csv \
     python tools/time_kernels.py $cfg > /dev/null 2>&1
  python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launch_$cfg.csv')) if len(r)>5 and r[0].isdigit()]
d=collections.defaultdict(list)
for r in rows:
    d[r[4].split('(')[0][-40:]].append(float(r[-1]))
for k,v in d.items():
    v=v[len(v)//2:]
    print('$cfg', k, 'n=%d'%len(v), 'mean %.1f'%(sum(v)/len(v)), r[-2])
PY
done
