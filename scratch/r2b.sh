#!/bin/bash
# round-2 evidence: tests + bench + reference arm + launch list + full capture of the top kernels (1 GPU)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gpu_tests.txt 2>&1; echo tests rc=$?
tail -3 gpurun_out/r02b_gpu_tests.txt
timeout 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02b_bench_reference.json 2>> gpurun_out/r02b_bench.err; echo ref rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02b_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_ncu_bench.log 2>&1; echo ncu rc=$?
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"head_bwd_adam_tc2|adam_stream|head_stats_tc|HeadDwAdam|GruBptt|GruStep|HeadCmaxFlat|HeadFwd" --launch-skip 150 -c 40 -o gpurun_out/r02b_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_ncu_full.log 2>&1; echo ncufull rc=$?
ls -la gpurun_out/r02b_full.ncu-rep
