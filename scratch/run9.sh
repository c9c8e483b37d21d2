#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r10_whatif.log; : > $L
echo "== baseline" >> $L; python scratch/prof_ct.py 32768 2>&1 | grep "^B " >> $L
echo "== primal x2" >> $L; SBM_KERNEL_DEFINES="-DSBM_CT_PRIMAL_REPEAT=2" python scratch/prof_ct.py 32768 2>&1 | grep "^B " >> $L
echo "== primal x3" >> $L; SBM_KERNEL_DEFINES="-DSBM_CT_PRIMAL_REPEAT=3" python scratch/prof_ct.py 32768 2>&1 | grep "^B " >> $L
echo "== tol 1e-6" >> $L; python scratch/prof_ct.py 32768 0,0,0 1e-6 2>&1 | grep "^B " >> $L
echo "== tol 1e-8" >> $L; python scratch/prof_ct.py 32768 0,0,0 1e-8 2>&1 | grep "^B " >> $L
cat $L
