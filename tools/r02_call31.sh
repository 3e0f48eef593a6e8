#!/bin/bash
timeout 900 python -m pytest tests/test_rpcfit.py -m gpu -x -q -s > gpurun_out/r02_pytest_rpcfit.log 2>&1; echo "pytest rpcfit exit $?"; grep -E "held-out|passed|failed|Error" gpurun_out/r02_pytest_rpcfit.log | cut -c1-400
bash tools/campaign_r02_n1.sh
