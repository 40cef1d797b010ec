#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2l_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2l_pytest_multi.log
SFM_BENCH_TRACE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2l_bench_${N}gpu.json 2> gpurun_out/r2l_bench_${N}gpu.err; echo "bench rc=$?"
grep "bench rank 0" gpurun_out/r2l_bench_${N}gpu.err | tail -4
python - <<PY
import json
l=json.loads(open('gpurun_out/r2l_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:l.get(k) for k in ('value','ms_per_step','n_gpus','loss_allreduce_check')})
print('e2e', l['e2e']['value'])
for name, ent in (l.get('strong_scaling') or {}).items():
    print(name, ent.get('single_gpu_us_per_step'))
    for k,v in ent.items():
        if isinstance(v, dict): print('   ', k, round(v['us_per_step'],1), round(v['efficiency_vs_single_gpu'],3), v.get('reduced_equals_world_x_local'))
PY
