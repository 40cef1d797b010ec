#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2e_pytest_gpu.log
timeout 600 python tools/exp_variants.py cfg1 cfg2 cfg4 cfg5 -- "" "SFM_SM_FRAC=50" "SFM_SM_FRAC=100" "SFM_SM_FRAC=30" 2>&1 | tee gpurun_out/r2e_time.log
bash tools/exp_launches.sh cfg2 cfg4 cfg5 2>&1 | grep -v march
