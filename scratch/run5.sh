#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r6_sweep.log; : > $L
for c in 2 3; do for sh in 1 2 3 4 6; do
echo "== C=$c shares=$sh" >> $L; SBM_CT_SHARES=$sh python scratch/prof_ct.py 32768 0,$c,2 2>&1 | grep "^B " >> $L
done; done
cat $L
