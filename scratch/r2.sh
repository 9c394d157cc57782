#!/bin/bash
# round-2 evidence run: tests, bench, launch list
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.txt 2>&1; echo tests rc=$?
tail -3 gpurun_out/r02_gpu_tests.txt
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo bench rc=$?
tail -3 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err; echo ref rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1; echo ncu rc=$?
