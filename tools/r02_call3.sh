#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_1m_a.csv python tools/profile_iter.py 1m 4 > gpurun_out/prof_iter_a.log 2>&1; echo "ncu list exit $?"; tail -3 gpurun_out/prof_iter_a.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_pt_assemble|k_pt_schur|k_pt_jvp1|k_pt_backsub|k_pt_reduce' -s 12 -c 8 -o gpurun_out/r02_full_a -f python tools/profile_iter.py 1m 3 > gpurun_out/ncu_full_a.log 2>&1; echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full_a.log
ls -la gpurun_out/r02_full_a.ncu-rep
