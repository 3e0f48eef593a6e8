#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/pt_check.py > gpurun_out/r02_pt_check.log 2>&1; echo "pt_check exit $?"; tail -9 gpurun_out/r02_pt_check.log | cut -c1-420
timeout 300 python tools/profile_iter.py 1m 8 > gpurun_out/prof_iter_e.log 2>&1; tail -1 gpurun_out/prof_iter_e.log
timeout 300 python tools/profile_iter.py cfg2 8 > gpurun_out/prof_iter_e2.log 2>&1; tail -1 gpurun_out/prof_iter_e2.log
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_pt_schur|k_pt_jvp1|k_pt_backsub|k_pt_assemble' -s 5 -c 8 -o gpurun_out/r02_full_e -f python tools/profile_iter.py 1m 3 > gpurun_out/ncu_full_e.log 2>&1; echo "ncu full exit $?"
