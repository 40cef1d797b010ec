#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
timeout 600 python tools/exp_variants.py cfg2 cfg5 -- "" 2>&1 | tee gpurun_out/r2n_time.log
