#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_pt_schur' -c 1 -o gpurun_out/r02_full_k3 -f python tools/profile_iter.py 1m 3 > gpurun_out/ncu_full_k3.log 2>&1; echo "ncu k3 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_pt_backsub|k_pt_jvp1' -c 2 -o gpurun_out/r02_full_k24 -f python tools/profile_iter.py 1m 3 > gpurun_out/ncu_full_k24.log 2>&1; echo "ncu k24 exit $?"
