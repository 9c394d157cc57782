#!/bin/bash
# usage: scratch/mg.sh N tag  -- N-GPU bench (default workload) [+ the NCCL equivalence tests when N == 2]
N=$1; tag=$2
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu or torchrun or regrowth" > gpurun_out/${tag}_tests_2gpu.txt 2>&1
  tail -3 gpurun_out/${tag}_tests_2gpu.txt
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err
echo bench rc=$?
head -c 3000 gpurun_out/${tag}_bench_${N}gpu.json; tail -5 gpurun_out/${tag}_bench_${N}gpu.err
