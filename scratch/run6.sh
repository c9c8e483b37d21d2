#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export SBM_CT_SHARES=3
bash scratch/run_ncu.sh r1_ct_v2_c2s3 --shape 0,2,2
L=gpurun_out/r7_c1.log; : > $L
for sh in 1 2 3; do echo "== C=1 shares=$sh" >> $L; SBM_CT_SHARES=$sh python scratch/prof_ct.py 32768 0,1,2 2>&1 | grep "^B " >> $L; done
cat $L
