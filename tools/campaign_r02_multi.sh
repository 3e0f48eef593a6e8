#!/bin/bash
# Round-2 multi-GPU run on N GPUs: bench (default workload, weak scaling), distributed-vs-single checks, and at N = 8 BASELINE configs 3 and 4.
N=$1
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29531 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N exit $?"
if [[ " $* " == *" check "* ]]; then
  run 29533 tools/dist_check.py --tight > gpurun_out/r02_dist_check_n$N.log 2>&1; echo "dist_check N=$N exit $?"; grep -v "Warn\|^\*\|OMP_NUM" gpurun_out/r02_dist_check_n$N.log | tail -7 | cut -c1-260
fi
if [[ " $* " == *" cfg3 "* ]]; then
  run 29532 bench.py --gpus $N --steps 10 --warmup 3 --workload cfg3 > gpurun_out/r02_bench_cfg3_n$N.json 2> gpurun_out/r02_bench_cfg3_n$N.err; echo "bench cfg3 N=$N exit $?"
fi
if [[ " $* " == *" cfg4 "* ]]; then
  run 29534 bench.py --gpus $N --steps 5 --warmup 2 --workload cfg4 > gpurun_out/r02_bench_cfg4_n$N.json 2> gpurun_out/r02_bench_cfg4_n$N.err; echo "bench cfg4 N=$N exit $?"; tail -3 gpurun_out/r02_bench_cfg4_n$N.err | cut -c1-300
fi
for f in gpurun_out/r02_bench_n$N.json gpurun_out/r02_bench_cfg3_n$N.json gpurun_out/r02_bench_cfg4_n$N.json; do
  [ -s $f ] && python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['n_gpus'], d['config']['n_obs'], d['config']['engine'], 'value %.4g' % d['value'], 'ms/it %.4f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'wall %.3f' % d['e2e']['wall_s'], d['e2e']['iterations'], d['e2e'].get('wall_breakdown_s')); print({k: round(v,4) for k,v in d['phases_ms_per_iteration'].items()})"
done
exit 0
