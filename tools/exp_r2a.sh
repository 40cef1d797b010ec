#!/bin/bash
# r2a: new parity tests at cfg4 / cfg5 shapes + first knob sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cfg4_cfg5" > gpurun_out/r2a_pytest_new.log 2>&1; echo "pytest(new) rc=$?"
tail -3 gpurun_out/r2a_pytest_new.log
timeout 900 python tools/exp_variants.py cfg2 cfg4 cfg5 -- "" "SFM_LIFO=1" > gpurun_out/r2a_lifo.log 2>&1
cat gpurun_out/r2a_lifo.log
timeout 600 python tools/exp_variants.py cfg4 -- "SFM_PF=1480" "SFM_PF=2960" "SFM_PF=5920" "SFM_LIFO=1 SFM_PF=2960" "SFM_HSEG=4" "SFM_HSEG=16" "SFM_HSEG=32" > gpurun_out/r2a_cfg4.log 2>&1
cat gpurun_out/r2a_cfg4.log
timeout 600 python tools/exp_variants.py cfg5 -- "SFM_SSIM_NW=2" "SFM_SSIM_NW=2 SFM_LIFO=1" "SFM_HSEG=32" "SFM_HSEG=43" > gpurun_out/r2a_cfg5.log 2>&1
cat gpurun_out/r2a_cfg5.log
for v in minb16 minb24; do
  echo "== $v"
  SFM_LIB_PATH=$PWD/sfm_learner_chainer_b200/variants/lib_$v.so timeout 300 python tools/exp_variants.py cfg4 -- "" "SFM_HSEG=16" 2>&1 | tee gpurun_out/r2a_$v.log
done
