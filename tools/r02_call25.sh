#!/bin/bash
mkdir -p gpurun_out
SBA_TIMING=1 timeout 300 python tools/e2e_time.py 1m > gpurun_out/r02_e2e_time.log 2>&1; tail -8 gpurun_out/r02_e2e_time.log
timeout 900 python -m pytest tests/test_gpu_pcg.py -m gpu -x -q > gpurun_out/r02_pytest_pcg.log 2>&1; echo "pytest pcg exit $?"; tail -12 gpurun_out/r02_pytest_pcg.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -5 gpurun_out/r02_pytest_gpu.log | cut -c1-300
