#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "eval or topk or tie or benchmarked" 2>&1 | tail -4
timeout 300 python bench.py --workload eval --no-cpu-baseline > gpurun_out/e10_eval.json 2>gpurun_out/e10_eval.err; tail -2 gpurun_out/e10_eval.err
python -c "
import json; d=json.load(open('gpurun_out/e10_eval.json')); print('eval', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"HeadCmaxPair|chunk_score64" --launch-skip 4 -c 2 -o /tmp/pair python bench.py --workload eval --no-secondary --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_ncu_pair.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/pair.ncu-rep --page raw --csv > gpurun_out/r02d_ncu_full_pair_raw.csv 2>/dev/null
