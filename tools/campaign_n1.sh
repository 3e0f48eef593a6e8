#!/bin/bash
# Round-end single-GPU measurement run: bench line, reference arm, ncu launch list and one full-set capture per kernel.
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r01_bench_n1.json 2> gpurun_out/r01_bench_n1.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference_n1.json 2> gpurun_out/r01_bench_reference.err; echo "ref exit $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_launches_cfg2.csv python tools/profile_iter.py cfg2 4 > gpurun_out/prof_iter.log 2>&1; echo "ncu list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_schur|k_assemble|k_jvp|k_chol|k_point_prep|k_residual|k_backsub|k_scale_dots|k_build_t2|k_step' -s 30 -c 16 -o gpurun_out/r01_full -f python tools/profile_iter.py cfg2 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/r01_full.ncu-rep
