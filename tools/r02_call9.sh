#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/pt_check.py > gpurun_out/r02_pt_check.log 2>&1; echo "pt_check exit $?"; tail -2 gpurun_out/r02_pt_check.log | cut -c1-420
timeout 300 python tools/profile_iter.py 1m 8 > gpurun_out/prof_iter_g.log 2>&1; tail -1 gpurun_out/prof_iter_g.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02_pytest_gpu.log
