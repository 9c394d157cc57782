#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "cfg3_like or bidir" > gpurun_out/d2_tests.txt 2>&1
tail -15 gpurun_out/d2_tests.txt
REC_TIMELINE=1 REC_NO_OVERLAP=1 N_STEPS=46 timeout 300 python scratch/time_cfg3.py > gpurun_out/d2_cfg3.txt 2>&1
timeout 300 python scratch/time_cfg3.py >> gpurun_out/d2_cfg3.txt 2>&1
tail -40 gpurun_out/d2_cfg3.txt
