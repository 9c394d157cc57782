#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for cfg in "REC_DW_RESERVE=24" "REC_DW_RESERVE=48" "REC_DW_RESERVE=24 REC_SWEEP_LATE=1" "REC_DW_RESERVE=48 REC_SWEEP_LATE=1" "REC_DW_RESERVE=0" "REC_NO_OVERLAP=1"; do
  echo "== $cfg"
  env $cfg timeout 300 python scratch/time_cfg3.py 2>&1 | tail -1
done
