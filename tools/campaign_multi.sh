#!/bin/bash
# Round-end multi-GPU measurement run: bench.py on N GPUs (weak scaling, config 2 per GPU), optionally the config-3 shard
# per GPU and the distributed-vs-single consistency check.   usage: tools/campaign_multi.sh N [cfg3] [check]
N=$1
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29531 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r01_bench_n$N.json 2> gpurun_out/r01_bench_n$N.err; echo "bench cfg2 N=$N exit $?"
if [[ " $* " == *" cfg3 "* ]]; then
  run 29532 bench.py --gpus $N --steps 10 --warmup 3 --workload cfg3 > gpurun_out/r01_bench_cfg3_n$N.json 2> gpurun_out/r01_bench_cfg3_n$N.err; echo "bench cfg3 N=$N exit $?"
fi
if [[ " $* " == *" check "* ]]; then
  run 29533 tools/dist_check.py > gpurun_out/r01_dist_check_n$N.log 2>&1; echo "dist_check N=$N exit $?"; tail -5 gpurun_out/r01_dist_check_n$N.log
fi
for f in gpurun_out/r01_bench_n$N.json gpurun_out/r01_bench_cfg3_n$N.json; do
  [ -s $f ] && python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['n_gpus'], d['config']['n_obs'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['wall_s']); print(d['phases_ms_per_iteration'])"
done
