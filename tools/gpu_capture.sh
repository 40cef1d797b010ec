#!/bin/bash
# ncu --set full captures of the fused kernel per config (and the prologue at cfg5) for profiles/ncu_latest.json and the
# summaries.  usage: tools/gpu_capture.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
for cfg in cfg2 cfg4 cfg5; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:march -s 3 -c 1 -f \
      -o gpurun_out/${tag}_fused_$cfg python tools/time_kernels.py $cfg > gpurun_out/${tag}_ncu_$cfg.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:prep -s 3 -c 1 -f \
      -o gpurun_out/${tag}_prologue_cfg5 python tools/time_kernels.py cfg5 > gpurun_out/${tag}_ncu_prep.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
