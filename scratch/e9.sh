#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "wide or cfg3" 2>&1 | tail -4
timeout 300 python bench.py --workload cfg3 --no-cpu-baseline --no-secondary > gpurun_out/e9_cfg3.json 2>gpurun_out/e9_cfg3.err; tail -2 gpurun_out/e9_cfg3.err
python -c "
import json; d=json.load(open('gpurun_out/e9_cfg3.json')); print('cfg3', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']); [print('   ',k[:60], round(v['kernel_ms'],4), round(v['frac'],3)) for k,v in d['roofline']['kernels'].items()]"
