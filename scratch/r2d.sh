#!/bin/bash
# round-2 evidence, final session (1 GPU): tests, bench, reference arm, launch list, full captures exported as CSV
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
P=r02d
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${P}_gpu_tests.txt 2>&1; echo tests rc=$?
tail -3 gpurun_out/${P}_gpu_tests.txt
timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo bench rc=$?
timeout 300 python bench.py --workload eval70k --no-cpu-baseline > gpurun_out/${P}_bench_eval70k.json 2>> gpurun_out/${P}_bench.err; echo eval70k rc=$?
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${P}_bench_reference.json 2>> gpurun_out/${P}_bench.err; echo ref rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/${P}_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${P}_ncu_bench.log 2>&1; echo ncu rc=$?
cap() { # name workload regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$3" --launch-skip $4 -c $5 -o /tmp/$1 python bench.py --workload $2 --no-secondary --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${P}_ncu_$1.log 2>&1
  echo "ncu $1 rc=$?"
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/${P}_ncu_full_$1_raw.csv 2>/dev/null
  rm -f /tmp/$1.ncu-rep
}
cap cfg4 cfg4 "head_bwd_adam_tc2|adam_stream|head_stats_tc" 30 10
cap eval eval "HeadCmaxPair|chunk_select2|chunk_score64" 6 3
du -sh gpurun_out
