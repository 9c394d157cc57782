#!/bin/bash
# round-2 evidence (1 GPU): tests, bench, reference arm, launch list, full captures exported as CSV (reports stay on the box)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gpu_tests.txt 2>&1; echo tests rc=$?
tail -3 gpurun_out/r02c_gpu_tests.txt
timeout 900 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02c_bench_reference.json 2>> gpurun_out/r02c_bench.err; echo ref rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/r02c_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_ncu_bench.log 2>&1; echo ncu rc=$?
cap() { # name workload regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$3" --launch-skip $4 -c $5 -o /tmp/$1 python bench.py --workload $2 --no-secondary --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_ncu_$1.log 2>&1
  echo "ncu $1 rc=$?"
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r02c_ncu_full_$1_raw.csv 2>/dev/null
  rm -f /tmp/$1.ncu-rep
}
cap cfg4 cfg4 "head_bwd_adam_tc2|adam_stream|head_stats_tc" 30 10
cap cfg3 cfg3 "tck_kernel" 44 11
cap eval eval "HeadCmaxFlat|chunk_" 3 3
du -sh gpurun_out
