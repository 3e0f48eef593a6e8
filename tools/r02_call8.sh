#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/pt_check.py > gpurun_out/r02_pt_check.log 2>&1; echo "pt_check exit $?"; tail -3 gpurun_out/r02_pt_check.log | cut -c1-420
timeout 300 python tools/profile_iter.py 1m 8 > gpurun_out/prof_iter_f.log 2>&1; tail -1 gpurun_out/prof_iter_f.log
timeout 300 python tools/profile_iter.py cfg2 8 > gpurun_out/prof_iter_f2.log 2>&1; tail -1 gpurun_out/prof_iter_f2.log
timeout 600 python bench.py --workload 1m --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_pt_1m.json 2> gpurun_out/r02_pt_1m.err; echo "1m exit $?"; tail -3 gpurun_out/r02_pt_1m.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_pt_1m.json"))
print(d["config"]["n_obs"], "ms/it", d["ms_per_step"], "phases", d["phases_ms_per_iteration"], "e2e", d["e2e"]["value"], d["e2e"]["wall_s"], d["e2e"]["wall_breakdown_s"], d["e2e"]["iterations"], "jac ms", d["jacobian_pass_ms"], "launches", d["gpu_launches"])
PY
