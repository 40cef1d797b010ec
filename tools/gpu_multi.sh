#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): the N-GPU test and the sharded bench line.  usage: tools/gpu_multi.sh <tag> <N>
tag=${1:-r2}; N=${2:-2}
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/${tag}_pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest_multi_${N}gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${tag}_bench_cfg2_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err; echo "bench rc=$?"
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 3 > gpurun_out/${tag}_bench_reference_${N}gpu.json 2>> gpurun_out/${tag}_bench_${N}gpu.err; echo "reference arm rc=$?"
python - <<PY
import json
l=json.loads(open('gpurun_out/${tag}_bench_cfg2_${N}gpu.json').read().strip().splitlines()[-1])
print({k:l.get(k) for k in ('value','ms_per_step','n_gpus','loss_allreduce_check')})
print('e2e', l['e2e']['value'])
for name, ent in (l.get('strong_scaling') or {}).items():
    print(name, ent.get('single_gpu_us_per_step'))
    for k,v in ent.items():
        if isinstance(v, dict) and 'us_per_step' in v: print('   ', k, round(v['us_per_step'],1), round(v['efficiency_vs_single_gpu'],3), v.get('reduced_equals_world_x_local'))
r=json.loads(open('gpurun_out/${tag}_bench_reference_${N}gpu.json').read().strip().splitlines()[-1])
print('reference arm', r['value'], r['config'] == l['config'])
PY
