#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r14_prof.log; : > $L
echo "== XW=1 shares=1" >> $L; SBM_CT_SHARES=1 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== XW=1 shares=3 primal x2" >> $L; SBM_KERNEL_DEFINES="-DSBM_CT_PRIMAL_REPEAT=2" python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== C=3 XW=0 (8 warps)" >> $L; SBM_CT_XW=0 python scratch/prof_ct.py 65536 0,3,2 2>&1 | grep "^B " >> $L
cat $L
