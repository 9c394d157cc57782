#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py -m gpu -x -q -k "not real_shape" > gpurun_out/d5_wide.txt 2>&1
tail -4 gpurun_out/d5_wide.txt
REC_TIMELINE=1 REC_NO_OVERLAP=1 N_STEPS=46 timeout 300 python scratch/time_cfg3.py > gpurun_out/d5_cfg3.txt 2>&1
timeout 300 python scratch/time_cfg3.py >> gpurun_out/d5_cfg3.txt 2>&1
tail -2 gpurun_out/d5_cfg3.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/d5_tests.txt 2>&1
tail -6 gpurun_out/d5_tests.txt
