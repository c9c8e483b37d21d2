#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r22_sweep.log; : > $L
echo "== default shares=1" >> $L; python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== shares=2" >> $L; SBM_CT_SHARES=2 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== shares=3" >> $L; SBM_CT_SHARES=3 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== primal x2" >> $L; SBM_KERNEL_DEFINES="-DSBM_CT_PRIMAL_REPEAT=2" python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== C=3 XW=0 tmem (8 warps)" >> $L; SBM_CT_XW=0 SBM_CT_SHARES=3 python scratch/prof_ct.py 65536 0,3,2 2>&1 | grep "^B " >> $L
cat $L
