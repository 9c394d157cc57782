#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sharded_r2.py -q > gpurun_out/f2_tests_new.txt 2>&1; echo new rc=$?
tail -40 gpurun_out/f2_tests_new.txt
