#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2f_pytest_gpu.log
timeout 600 python tools/exp_variants.py cfg1 cfg4 -- "" "SFM_HSEG=16" "SFM_HSEG=4" 2>&1 | tee gpurun_out/r2f_time.log
for v in p12 p20 s1m20 s1m24; do
  echo "== $v"
  SFM_LIB_PATH=$PWD/sfm_learner_chainer_b200/variants/lib_$v.so timeout 300 python tools/exp_variants.py cfg1 cfg4 -- "" "SFM_HSEG=16" 2>&1 | tee gpurun_out/r2f_$v.log
done
