#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
REC_TIMELINE=1 N_STEPS=46 timeout 300 python scratch/time_cfg3.py > gpurun_out/d6_cfg3_ov.txt 2>&1
grep timeline gpurun_out/d6_cfg3_ov.txt | tail -42
