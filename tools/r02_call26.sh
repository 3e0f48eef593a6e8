#!/bin/bash
# 2 GPUs: pattern engine distributed parity + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pcg.py tests/test_gpu_parity_at_size.py -m gpu -x -q -k "pcg or multi_gpu" > gpurun_out/r02_pytest_n2.log 2>&1; echo "pytest n2 exit $?"; tail -8 gpurun_out/r02_pytest_n2.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dist_check.py > gpurun_out/r02_dist_check_n2.log 2>&1; echo "dist_check exit $?"; tail -6 gpurun_out/r02_dist_check_n2.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 exit $?"; tail -3 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n2.json"))
print(d["n_gpus"], d["config"]["n_obs"], d["config"]["engine"], "value", d["value"], "ms/it", d["ms_per_step"], "phases", d["phases_ms_per_iteration"])
print("e2e", d["e2e"]["value"], d["e2e"]["wall_s"], d["e2e"]["iterations"])
PY
