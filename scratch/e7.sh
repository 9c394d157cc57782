#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/e7_tests.txt 2>&1; tail -4 gpurun_out/e7_tests.txt
for w in eval eval70k; do
timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/e7_$w.json 2>gpurun_out/e7_$w.err; tail -2 gpurun_out/e7_$w.err
python -c "
import json; d=json.load(open('gpurun_out/e7_$w.json')); print('$w', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])"
done
