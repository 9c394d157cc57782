#!/bin/bash
# usage: scratch/mg4.sh N -- N-GPU runs of the final round-2 session: NCCL equivalence (N == 2, row-sharded table on), bench with
# the replicated table (+ evaluation sharded by sessions), bench with the row-sharded table (both trunks at N == 4)
N=$1
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
P=r02e
run() { # tag, env..., extra args after --
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus $N --steps 100 --warmup 20 "$@" > gpurun_out/${P}_bench_${N}gpu_${tag}.out 2> gpurun_out/${P}_bench_${N}gpu_${tag}.err
  echo "$tag rc=$?"
  grep "^{" gpurun_out/${P}_bench_${N}gpu_${tag}.out > gpurun_out/${P}_bench_${N}gpu_${tag}.json
  python - <<EOF
import json
try:
    d=json.load(open('gpurun_out/${P}_bench_${N}gpu_${tag}.json'))
    ev=d.get('secondary',{}).get('eval',{})
    print('$tag', d['n_gpus'], 'sessions/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'eval', ev.get('value'), ev.get('ms_per_step'), ev.get('ranks_agree_on_metrics'), d['replicated_params_bit_identical_across_ranks'], d['config']['parallelism'][:90])
except Exception as ex: print('$tag', 'no line', ex)
EOF
}
if [ "$N" = "2" ]; then
  REC_SHARD_EMBEDDING=1 timeout 600 python -m pytest tests -m gpu -x -q -k "multi_gpu" > gpurun_out/${P}_tests_2gpu.txt 2>&1
  tail -3 gpurun_out/${P}_tests_2gpu.txt
fi
run replicated_table REC_SHARD_EMBEDDING=0 --
run row_sharded_table REC_SHARD_EMBEDDING=1 -- --no-secondary
if [ "$N" = "4" ]; then
  run row_sharded_table_dp_trunk REC_SHARD_EMBEDDING=1 REC_DP_TRUNK=1 -- --no-secondary
fi
for f in gpurun_out/${P}_bench_${N}gpu_*.err; do tail -n 2 $f | cut -c1-300; done
