#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sarm" > gpurun_out/d13_tests.txt 2>&1
tail -25 gpurun_out/d13_tests.txt
