#!/bin/bash
# chunked supervised-head backward with the dh read-out moved behind the next publish: correctness + timing at B = 1024
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "large_batch_chunked or tensor_core_heads_agree or head_gradient_wrt_state" > gpurun_out/f10_tests.txt 2>&1; echo tests rc=$?
tail -3 gpurun_out/f10_tests.txt
for B in 1024; do
  timeout 200 python bench.py --workload cfg4 --batch $B --no-secondary --no-cpu-baseline --steps 40 --warmup 5 > gpurun_out/f10_bench_cfg4_B$B.json 2> gpurun_out/f10_bench_cfg4_B$B.err
  echo "B=$B rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/f10_bench_cfg4_B$B.json')); r=d['roofline']
print($B, round(d['value']), d['ms_per_step'], {k[:28]: round(v['kernel_ms'],3) for k,v in r['kernels'].items()})" || tail -n 5 gpurun_out/f10_bench_cfg4_B$B.err | cut -c1-300
done
