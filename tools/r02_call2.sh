#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/pt_check.py > gpurun_out/r02_pt_check.log 2>&1; echo "pt_check exit $?"; tail -15 gpurun_out/r02_pt_check.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --workload 1m --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_pt_1m.json 2> gpurun_out/r02_pt_1m.err; echo "1m exit $?"; tail -3 gpurun_out/r02_pt_1m.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_pt_1m.json",):
    try:
        d = json.load(open(f))
        print(f, d["config"]["n_obs"], "ms/it", d["ms_per_step"], "phases", d["phases_ms_per_iteration"], "e2e", d["e2e"]["value"], d["e2e"]["wall_s"], d["e2e"]["wall_breakdown_s"], "jac ms", d["jacobian_pass_ms"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "failed", e)
PY
