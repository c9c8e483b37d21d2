#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r17_whatif.log; : > $L
echo "== default" >> $L; SBM_CT_HEAD_SHARES=2 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== G loaded twice" >> $L; SBM_CT_HEAD_SHARES=2 SBM_KERNEL_DEFINES="-DSBM_CT_WHATIF_G_TWICE=1" python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== primal x2 (shares=1)" >> $L; SBM_CT_HEAD_SHARES=2 SBM_KERNEL_DEFINES="-DSBM_CT_PRIMAL_REPEAT=2" python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
cat $L
