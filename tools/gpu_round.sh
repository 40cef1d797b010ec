#!/bin/bash
# One gpurun call: GPU parity tests, bench line (driver flags and defaults), reference arm, per-config timings, launch list,
# ncu full captures.  usage: tools/gpu_round.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${tag}_clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/${tag}_bench_cfg2_driverflags.json 2> gpurun_out/${tag}_bench.err; echo "bench(driver flags) rc=$?"
timeout 900 python bench.py > gpurun_out/${tag}_bench_cfg2.json 2>> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench_cfg2.json | cut -c1-600
timeout 300 python bench.py --impl reference --steps 5 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench_reference.json | cut -c1-300
timeout 300 python tools/time_graph.py cfg1 cfg2 cfg4 cfg5 > gpurun_out/${tag}_time_kernels.log 2>&1
cat gpurun_out/${tag}_time_kernels.log
timeout 300 python tools/time_cfg3.py 2>/dev/null | grep cfg3 > gpurun_out/${tag}_time_cfg3.log
cat gpurun_out/${tag}_time_cfg3.log
kill $SMI
# launch list of the bench command itself (graph replay is opaque to ncu's per-kernel list -> direct calls)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg2.csv \
    python bench.py --steps 20 --warmup 3 --no-other --no-cpu --no-graph > gpurun_out/${tag}_b_ncu.log 2>&1
for cfg in cfg2 cfg4 cfg5; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:march -s 3 -c 1 -f \
      -o gpurun_out/${tag}_fused_$cfg python tools/time_kernels.py $cfg > gpurun_out/${tag}_ncu_$cfg.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:prep -s 3 -c 1 -f \
      -o gpurun_out/${tag}_prologue_cfg5 python tools/time_kernels.py cfg5 > gpurun_out/${tag}_ncu_prep.log 2>&1
ls -la gpurun_out | tail -12
