#!/bin/bash
# usage: scratch/mg2.sh N  -- N-GPU default bench (+ reference arm line, + NCCL equivalence tests at N == 2)
N=$1
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu or torchrun or regrowth" > gpurun_out/r02d_tests_2gpu.txt 2>&1
  tail -3 gpurun_out/r02d_tests_2gpu.txt
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r02d_bench_${N}gpu.out 2> gpurun_out/r02d_bench_${N}gpu.err
echo bench rc=$?
grep "^{" gpurun_out/r02d_bench_${N}gpu.out > gpurun_out/r02d_bench_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r02d_bench_${N}gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['secondary']['eval']['value'], d['replicated_params_bit_identical_across_ranks'], d['clocks'])"
