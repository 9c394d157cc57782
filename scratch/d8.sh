#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python scratch/trace_gru.py > gpurun_out/d8_trace.txt 2>&1
sed -n 10,20p gpurun_out/d8_trace.txt
for cfg in "REC_DW_RESERVE=40" "REC_DW_RESERVE=56" "REC_DW_RESERVE=32"; do
  echo "== $cfg"
  env $cfg timeout 300 python scratch/time_cfg3.py 2>&1 | tail -1
done
