#!/bin/bash
# Round-2 first GPU call: FP64 / shared-memory micro-benchmarks and the round-1 engine at the round-2 bench sizes.
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_peak tools/fp64_peak.cu && /tmp/fp64_peak > gpurun_out/r02_fp64_peak.json; echo "fp64 exit $?"; cat gpurun_out/r02_fp64_peak.json
timeout 600 python bench.py --workload 1m --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_base_1m.json 2> gpurun_out/r02_base_1m.err; echo "1m exit $?"
timeout 900 python bench.py --workload cfg3full --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_base_cfg3full.json 2> gpurun_out/r02_base_cfg3full.err; echo "cfg3full exit $?"
tail -c 600 gpurun_out/r02_base_1m.err gpurun_out/r02_base_cfg3full.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_base_1m.json", "gpurun_out/r02_base_cfg3full.json"):
    try:
        d = json.load(open(f))
        print(f, d["config"]["n_obs"], "ms/it", d["ms_per_step"], "phases", d["phases_ms_per_iteration"], "e2e", d["e2e"]["value"], d["e2e"]["wall_s"], "jac ms", d["jacobian_pass_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
