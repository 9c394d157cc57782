#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "ties or vocab_sharded_phases or evaluate" > gpurun_out/d14_tests.txt 2>&1
tail -12 gpurun_out/d14_tests.txt
