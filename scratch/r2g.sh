#!/bin/bash
# last verification of the round: full GPU suite + the default workload alone (cfg4, B = 256) after the backward-kernel change
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
P=r02g
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${P}_gpu_tests.txt 2>&1; echo tests rc=$?
tail -3 gpurun_out/${P}_gpu_tests.txt
timeout 200 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/${P}_bench_cfg4_only.json 2> gpurun_out/${P}_bench.err; echo bench rc=$?
python -c "
import json
d=json.load(open('gpurun_out/${P}_bench_cfg4_only.json')); r=d['roofline']
print(round(d['value']), d['ms_per_step'], r['frac'], {k[:28]: round(v['kernel_ms'],4) for k,v in r['kernels'].items()}, d['clocks'])"
