#!/bin/bash
mkdir -p gpurun_out
for v in a b c d; do
  echo "== $v"
  SFM_LIB_PATH=$PWD/sfm_learner_chainer_b200/variants/lib_$v.so timeout 300 python tools/exp_variants.py cfg1 cfg4 -- "" "SFM_HSEG=16" "SFM_HSEG=4" 2>&1 | tee gpurun_out/r2g_$v.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march -s 3 -c 1 -f -o gpurun_out/r2g_fused_cfg4 python tools/time_kernels.py cfg4 > gpurun_out/r2g_ncu_cfg4.log 2>&1
