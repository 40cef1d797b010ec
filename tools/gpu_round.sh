#!/bin/bash
# One gpurun call: GPU parity tests, bench line, launch list, ncu full captures of the fused kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"
cat gpurun_out/bench_cfg2.json
timeout 300 python tools/time_kernels.py cfg1 cfg2 cfg4 cfg5 > gpurun_out/time_kernels.log 2>&1
cat gpurun_out/time_kernels.log
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv \
    python bench.py --steps 20 --warmup 3 --no-other --no-cpu --no-graph > gpurun_out/b_ncu.log 2>&1
for cfg in cfg2 cfg4 cfg5; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:march -s 3 -c 1 -f \
      -o gpurun_out/fused_$cfg python tools/time_kernels.py $cfg > gpurun_out/ncu_$cfg.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:prep -s 3 -c 1 -f \
      -o gpurun_out/prep_cfg5 python tools/time_kernels.py cfg5 > gpurun_out/ncu_prep.log 2>&1
ls -la gpurun_out
