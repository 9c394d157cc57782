#!/bin/bash
# single-GPU runs of cfg4 at B in {1024, 4096} (the tensor-bound regime)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for B in 1024 4096; do
  timeout 200 python bench.py --workload cfg4 --batch $B --no-secondary --no-cpu-baseline --steps 40 --warmup 5 > gpurun_out/r02f_bench_cfg4_B$B.json 2> gpurun_out/r02f_bench_cfg4_B$B.err
  echo "B=$B rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r02f_bench_cfg4_B$B.json')); r=d['roofline']
print($B, round(d['value']), d['ms_per_step'], 'hbm frac', round(r['frac'],3), 'tensor', r['tensor_side'], {k[:28]: round(v['kernel_ms'],3) for k,v in r['kernels'].items()}, d['clocks'])" || tail -n 5 gpurun_out/r02f_bench_cfg4_B$B.err | cut -c1-300
done
