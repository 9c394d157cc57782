#!/bin/bash
# A/B of the streaming-Adam footprint (CTAs per SM x unroll) on the cfg2 train step
for cfg in "1 2" "2 2" "3 2"; do
  set -- $cfg
  REC_SWEEP_CTAS=$1 REC_SWEEP_UNROLL=$2 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('ctas $1 unroll $2:', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['ms_per_step'],4), ' q-sweep', [round(v['kernel_ms'],4) for v in k.values()])"
done
