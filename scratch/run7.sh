#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r8_sweep.log; : > $L
run() { echo "== C=$1 shares=$2 warps=$3" >> $L; SBM_CT_SHARES=$2 SBM_CT_SHARE_WARPS=$3 python scratch/prof_ct.py 32768 0,$1,2 2>&1 | grep "^B " >> $L; }
run 2 2 ""
run 2 3 ""
run 2 4 ""
run 2 6 ""
run 2 3 "3,7,10"
run 2 2 "3,7"
run 2 4 "3,7,10,9"
run 3 2 ""
run 3 3 ""
run 3 2 "3,7"
cat $L
