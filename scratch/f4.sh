#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 400 python scratch/eval_trained.py 1024 243 > gpurun_out/f4_eval_trained.txt 2>&1
grep "REC_EVAL_TRACE\|total\|Error" gpurun_out/f4_eval_trained.txt | cut -c1-300
