#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2i_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2i_pytest_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err; echo "bench rc=$?"
tail -5 gpurun_out/r2i_bench_2gpu.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2i_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:l.get(k) for k in ('value','ms_per_step','n_gpus','loss_allreduce_check')})
print(l['timing']['parallelism'])
print('e2e', l['e2e']['value'])
for name, ent in (l.get('strong_scaling') or {}).items():
    print(name, ent.get('single_gpu_us_per_step'))
    for k,v in ent.items():
        if isinstance(v, dict): print('   ', k, round(v['us_per_step'],1), round(v['efficiency_vs_single_gpu'],3), v.get('reduced_equals_world_x_local'), v.get('collective'))
PY
