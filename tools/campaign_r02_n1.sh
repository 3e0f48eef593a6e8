#!/bin/bash
# Round-2 single-GPU measurement run: bench lines (default 1m + secondary cfg2, rpc workload), reference arms, ncu launch list of
# the bench command itself and one full-set capture per kernel of the 1m workload.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference.err; echo "ref exit $?"
timeout 600 python bench.py --workload rpc --steps 20 > gpurun_out/r02_bench_rpc.json 2> gpurun_out/r02_bench_rpc.err; echo "rpc exit $?"; tail -2 gpurun_out/r02_bench_rpc.err | cut -c1-300
timeout 600 python bench.py --workload 5m --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_5m.json 2> gpurun_out/r02_bench_5m.err; echo "5m exit $?"
timeout 600 python bench.py --workload cfg3full --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_cfg3full.json 2> gpurun_out/r02_bench_cfg3full.err; echo "cfg3full exit $?"
timeout 600 python bench.py --workload rpcba --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_rpcba.json 2> gpurun_out/r02_bench_rpcba.err; echo "rpcba exit $?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_under_ncu.json 2> /dev/null; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_pt_|k_chol_fused' -s 8 -c 9 -o gpurun_out/r02_full_1m -f python tools/profile_iter.py 1m 3 > gpurun_out/ncu_full_1m.log 2>&1; echo "ncu full exit $?"
(timeout 600 compute-sanitizer --tool memcheck python tools/profile_iter.py small 3; SBA_ENGINE=generic timeout 600 compute-sanitizer --tool memcheck python tools/profile_iter.py small 3; timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_triangulate.py -m gpu -x -q -k "edge or golden or oracle-perspective-10") > gpurun_out/r02_memcheck_raw.log 2>&1; grep -h "ERROR SUMMARY\|passed\|failed\|engine\|Error" gpurun_out/r02_memcheck_raw.log | cut -c1-200 > gpurun_out/r02_memcheck.log; cat gpurun_out/r02_memcheck.log
SBA_CHOL_CLK=1 python tools/chol_time.py 30 60 90 120 300 1800 > gpurun_out/r02_chol_time.log 2>&1; tail -14 gpurun_out/r02_chol_time.log | cut -c1-200
python - <<'PY'
import json
for f in ("r02_bench_n1", "r02_bench_5m", "r02_bench_cfg3full"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, d["config"]["n_obs"], d["config"]["engine"], "value %.4g" % d["value"], "ms/it %.4f" % d["ms_per_step"], {k: round(v, 4) for k, v in d["phases_ms_per_iteration"].items()}, "e2e %.4g" % d["e2e"]["value"], "jac %.4f" % d["jacobian_pass_ms"])
    except Exception as e:
        print(f, "failed", e)
try:
    d = json.load(open("gpurun_out/r02_bench_rpc.json"))
    print({k: ("%.3g" % v["value"], v["unit"], "cpu %.3g" % v.get("cpu_baseline", 0)) for k, v in d["operations"].items()}, d["roofline"]["frac"], d["roofline"]["fp64"]["frac"])
except Exception as e:
    print("rpc failed", e)
PY
