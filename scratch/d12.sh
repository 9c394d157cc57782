#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/d12_launches.csv python bench.py --workload eval --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/d12_ncu.log 2>&1
tail -1 gpurun_out/d12_ncu.log | cut -c1-200
