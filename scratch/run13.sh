#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r15_rows.log; : > $L
for r in 1 2 4 8; do
echo "== rows=$r shares=3" >> $L; SBM_KERNEL_DEFINES="-DSBM_CT_ROWS=$r" python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
done
echo "== rows=2 shares=1" >> $L; SBM_CT_SHARES=1 SBM_KERNEL_DEFINES="-DSBM_CT_ROWS=2" python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== rows=4 shares=1" >> $L; SBM_CT_SHARES=1 SBM_KERNEL_DEFINES="-DSBM_CT_ROWS=4" python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
cat $L
