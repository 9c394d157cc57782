#!/bin/bash
# N-GPU default bench line of the final round-2 state (row-sharded embedding sweep + evaluation sharded by sessions are the defaults)
N=$1
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r02f_bench_${N}gpu.out 2> gpurun_out/r02f_bench_${N}gpu.err
echo rc=$?
grep "^{" gpurun_out/r02f_bench_${N}gpu.out > gpurun_out/r02f_bench_${N}gpu.json
python -c "
import json
d=json.load(open('gpurun_out/r02f_bench_${N}gpu.json')); ev=d['secondary']['eval']
print(d['n_gpus'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'eval', round(ev['value']), ev['ms_per_step'], ev['sweeps_ms'], ev['ranks_agree_on_metrics'], d['replicated_params_bit_identical_across_ranks'], d['clocks'])
print(d['config']['parallelism'])"
tail -n 3 gpurun_out/r02f_bench_${N}gpu.err | cut -c1-300
