#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/pt_check.py > gpurun_out/r02_pt_check.log 2>&1; echo "pt_check exit $?"; tail -12 gpurun_out/r02_pt_check.log
timeout 300 python tools/profile_iter.py 1m 8 > gpurun_out/prof_iter_b.log 2>&1; tail -2 gpurun_out/prof_iter_b.log
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_pt_assemble|k_pt_schur|k_pt_jvp1|k_pt_backsub' -s 6 -c 6 -o gpurun_out/r02_full_b -f python tools/profile_iter.py 1m 3 > gpurun_out/ncu_full_b.log 2>&1; echo "ncu full exit $?"
