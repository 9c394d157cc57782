#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py -m gpu -x -q -k "not real_shape" > gpurun_out/d3_tests.txt 2>&1
tail -30 gpurun_out/d3_tests.txt
timeout 300 python -m pytest tests -m gpu -x -q -k "cfg3_like" > gpurun_out/d3_tests2.txt 2>&1
tail -5 gpurun_out/d3_tests2.txt
REC_TIMELINE=1 REC_NO_OVERLAP=1 N_STEPS=46 timeout 300 python scratch/time_cfg3.py > gpurun_out/d3_cfg3.txt 2>&1
timeout 300 python scratch/time_cfg3.py >> gpurun_out/d3_cfg3.txt 2>&1
tail -3 gpurun_out/d3_cfg3.txt
