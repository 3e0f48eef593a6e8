#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_at_size.py -m gpu -x -q > gpurun_out/r02_pytest_size.log 2>&1; echo "pytest size exit $?"; tail -15 gpurun_out/r02_pytest_size.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n1.json"))
print(d["config"]["n_obs"], d["config"]["engine"], "value", d["value"], "ms/it", d["ms_per_step"], "phases", d["phases_ms_per_iteration"])
print("e2e", d["e2e"]["value"], d["e2e"]["wall_s"], d["e2e"]["first_call_wall_s"], d["e2e"]["wall_breakdown_s"], d["e2e"]["iterations"], "jac ms", d["jacobian_pass_ms"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]); print("cpu", d.get("cpu_baseline")); print("secondary", d.get("secondary"))
PY
