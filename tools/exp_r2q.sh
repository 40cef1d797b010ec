#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "host" 2>&1 | tail -2
timeout 600 python tools/e2e_sweep.py cfg2 2>&1 | tee gpurun_out/r2q_e2e.log
