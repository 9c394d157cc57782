#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DIST_REGROW=1 REC_NO_DP_TRUNK=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29617 tests/dist_equivalence.py > gpurun_out/d1_regrow.txt 2>&1
echo regrow rc=$?
REC_TIMELINE=1 REC_NO_OVERLAP=1 timeout 200 python bench.py --workload cfg2 --steps 50 --warmup 5 --no-cpu-baseline --no-secondary 2>&1 | grep timeline | head -40 > gpurun_out/d1_tl_cfg2.txt
REC_TIMELINE=1 REC_NO_OVERLAP=1 timeout 200 python bench.py --workload cfg4 --steps 50 --warmup 5 --no-cpu-baseline --no-secondary 2>&1 | grep timeline | head -40 > gpurun_out/d1_tl_cfg4.txt
REC_TIMELINE=1 REC_NO_OVERLAP=1 N_STEPS=46 timeout 300 python scratch/time_cfg3.py > gpurun_out/d1_cfg3.txt 2>&1
timeout 300 python scratch/time_cfg3.py >> gpurun_out/d1_cfg3.txt 2>&1
tail -30 gpurun_out/d1_regrow.txt
