#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "evaluate_default_init" > gpurun_out/d11_tests.txt 2>&1
tail -5 gpurun_out/d11_tests.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tck_kernel|chunk_" -c 6 -o gpurun_out/d11_eval python bench.py --workload eval --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/d11_ncu.log 2>&1
tail -3 gpurun_out/d11_ncu.log
ls -la gpurun_out/d11_eval.ncu-rep
