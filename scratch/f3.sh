#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29655 scratch/eval_ng.py > gpurun_out/f3_eval_$1gpu.txt 2>&1
grep "REC_EVAL_TRACE\|total\|Error" gpurun_out/f3_eval_$1gpu.txt | cut -c1-400
