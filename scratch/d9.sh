#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "eval or topk or tie or native_training_loop or sharded_evaluation" > gpurun_out/d9_tests.txt 2>&1
tail -5 gpurun_out/d9_tests.txt
timeout 300 python bench.py --workload eval --no-cpu-baseline > gpurun_out/d9_eval.json 2>gpurun_out/d9_eval.err; tail -2 gpurun_out/d9_eval.err
python -c "
import json; d=json.load(open('gpurun_out/d9_eval.json')); print('eval', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
REC_NO_TCK_TOPK=1 timeout 300 python bench.py --workload eval --no-cpu-baseline > gpurun_out/d9_eval_old.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/d9_eval_old.json')); print('eval old', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'])"
timeout 300 python bench.py --workload eval70k --no-cpu-baseline > gpurun_out/d9_eval70k.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/d9_eval70k.json')); print('eval70k', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'])"
REC_NO_TCK_TOPK=1 timeout 300 python bench.py --workload eval70k --no-cpu-baseline > gpurun_out/d9_eval70k_old.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/d9_eval70k_old.json')); print('eval70k old', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'])"
