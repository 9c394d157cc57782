#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for B in 4096 512; do
  timeout 120 python bench.py --workload cfg4 --batch $B --no-secondary --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r02g_bench_cfg4_B$B.json 2> gpurun_out/r02g_bench_cfg4_B$B.err
  echo "B=$B rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r02g_bench_cfg4_B$B.json')); r=d['roofline']
print($B, round(d['value']), d['ms_per_step'], round(r['frac'],3), round(r['tensor_side']['frac'],3), {k[:28]: round(v['kernel_ms'],3) for k,v in r['kernels'].items()})"
done
