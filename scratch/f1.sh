#!/bin/bash
# new multi-rank tests on one GPU (virtual ranks + two gloo processes sharing cuda:0) + the sharded tests that existed
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded_r2.py -q > gpurun_out/f1_tests_new.txt 2>&1; echo new rc=$?
tail -30 gpurun_out/f1_tests_new.txt
