#!/bin/bash
# One-stop evidence run for profiles/ (run under gpurun on ONE GPU, after the plain commands have exited 0):
#   scratch/prof.sh <tag> [kernel-regex]
# writes gpurun_out/<tag>_bench_cfg2.json, <tag>_bench_cfg4.json, <tag>_launches_cfg2.csv and <tag>_full.ncu-rep;
# summarise the report here with
#   ncu -i gpurun_out/<tag>_full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,...
set -u
tag=${1:-x}
regex=${2:-"head_bwd_adam_tc2|head_stats_tc|adam_stream|gru_fwd64"}
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 20 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err || exit 1
python bench.py --workload cfg4 --steps 50 --warmup 5 > gpurun_out/${tag}_bench_cfg4.json 2>/dev/null || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/${tag}_launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$regex" --launch-skip 40 -c 10 \
  -o gpurun_out/${tag}_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu2.log 2>&1
ls -la gpurun_out/${tag}_*
