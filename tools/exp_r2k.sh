#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest_gpu.log
