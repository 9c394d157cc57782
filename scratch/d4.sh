#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py -m gpu -x -q -k "not real_shape" > gpurun_out/d4_tests.txt 2>&1
tail -4 gpurun_out/d4_tests.txt
REC_TIMELINE=1 REC_NO_OVERLAP=1 N_STEPS=46 timeout 300 python scratch/time_cfg3.py > gpurun_out/d4_cfg3.txt 2>&1
timeout 300 python scratch/time_cfg3.py >> gpurun_out/d4_cfg3.txt 2>&1
tail -2 gpurun_out/d4_cfg3.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/d4_launches.csv python scratch/time_cfg3.py > gpurun_out/d4_ncu.log 2>&1
tail -2 gpurun_out/d4_ncu.log
