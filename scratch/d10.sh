#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "eval or topk or tie or native_training_loop" > gpurun_out/d10_tests.txt 2>&1
tail -5 gpurun_out/d10_tests.txt
for w in eval eval70k; do
timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/d10_$w.json 2>gpurun_out/d10_$w.err; tail -2 gpurun_out/d10_$w.err
python -c "
import json; d=json.load(open('gpurun_out/d10_$w.json')); print('$w', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['metrics_sample'])"
done
REC_TIMELINE=1 timeout 300 python bench.py --workload eval --no-cpu-baseline 2>&1 | grep timeline | tail -12
