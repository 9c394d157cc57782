#!/bin/bash
N=$1
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
REC_EVAL_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus $N --steps 50 --warmup 20 > gpurun_out/f5_bench_${N}gpu.out 2> gpurun_out/f5_bench_${N}gpu.err
echo rc=$?
grep "REC_EVAL_TRACE" gpurun_out/f5_bench_${N}gpu.out | cut -c1-300
grep "^{" gpurun_out/f5_bench_${N}gpu.out | python -c "
import json,sys
d=json.loads(sys.stdin.read()); ev=d['secondary']['eval']; print(d['n_gpus'], d['value'], d['ms_per_step'], 'eval', ev['value'], ev['ms_per_step'])"
