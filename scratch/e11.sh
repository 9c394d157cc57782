#!/bin/bash
# experiment: part of the Q-head sweep next to the supervised-head kernel on disjoint SMs
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { # label env...
  l=$1; shift
  env "$@" timeout 200 python bench.py --workload cfg4 --no-secondary --no-cpu-baseline --steps 100 --warmup 10 > gpurun_out/e11.json 2>gpurun_out/e11.err || tail -3 gpurun_out/e11.err
  python -c "
import json; d=json.load(open('gpurun_out/e11.json')); print('$l', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
}
run base X=1
run early_nosplit REC_SWEEP_EARLY=1
run s25_g9_b120 REC_SWEEP_SPLIT=25 REC_SWEEP_GRID1=9 REC_BWD_CTAS=120
run s35_g13_b108 REC_SWEEP_SPLIT=35 REC_SWEEP_GRID1=13 REC_BWD_CTAS=108
run s20_g6_b130 REC_SWEEP_SPLIT=20 REC_SWEEP_GRID1=6 REC_BWD_CTAS=130
run s30_g9_b120_u4 REC_SWEEP_SPLIT=30 REC_SWEEP_GRID1=9 REC_BWD_CTAS=120 REC_SWEEP_UNROLL=4
run s40_g18_b92 REC_SWEEP_SPLIT=40 REC_SWEEP_GRID1=18 REC_BWD_CTAS=92
