#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2h_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print('roofline', {k:l['roofline'][k] for k in ('frac','kernel_us','traffic','traffic_source')})
print('step', l['roofline_step']['frac'])
print('e2e', {k:l['e2e'][k] for k in ('value','repeats','h2d_gbs','steps_per_repeat')}, l['e2e']['float_images']['value'], l['e2e']['synchronous']['value'])
print('others', {k:(round(v['ms_per_step']*1e3,1), round(v['fused_kernel_us'],1), round(v['step_frac_of_hbm_peak'],4)) for k,v in l['other_configs'].items()})
print('cpu', l['cpu_baseline'])
PY
timeout 300 python bench.py --impl reference --steps 5 > gpurun_out/r2h_ref.json 2>/dev/null; python -c "
import json;l=json.loads(open('gpurun_out/r2h_ref.json').read().strip().splitlines()[-1]);print(l['value'], l['cpu_baseline'])"
