#!/bin/bash
# exponent-evaluation variants of the pair kernel (REC_CMAX_EXP): where does the epilogue time go?
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in ${VARIANTS:-0 9 1 2 3 4}; do
REC_CMAX_EXP=$v timeout 300 python bench.py --workload eval --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/e2_$v.json 2>gpurun_out/e2_$v.err; tail -2 gpurun_out/e2_$v.err
python -c "
import json; d=json.load(open('gpurun_out/e2_$v.json')); print('variant $v', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
