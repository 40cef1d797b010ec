#!/bin/bash
# ncu full captures of the fused kernels for the given configs: tools/gpu_prof.sh <tag> cfg4 cfg5 ...
tag=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:march -s 3 -c 1 -f \
      -o gpurun_out/${tag}_fused_$cfg python tools/time_kernels.py $cfg > gpurun_out/${tag}_ncu_$cfg.log 2>&1
  tail -2 gpurun_out/${tag}_ncu_$cfg.log
done
ls -la gpurun_out | tail -8
