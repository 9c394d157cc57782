#!/bin/bash
# round-2 evidence, last session (1 GPU): full GPU test suite, default bench line, reference arm
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
P=r02f
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${P}_gpu_tests.txt 2>&1; echo tests rc=$?
tail -3 gpurun_out/${P}_gpu_tests.txt
timeout 600 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${P}_bench_reference.json 2>> gpurun_out/${P}_bench.err; echo ref rc=$?
python -c "
import json
d=json.load(open('gpurun_out/${P}_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], {k:(v.get('value'), v.get('ms_per_step')) for k,v in d['secondary'].items()})"
