#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/sweep.log; : > $L
for sh in 1 4 7 10; do echo "== shares=$sh" >> $L; SBM_CT_SHARES=$sh python tools/prof_ct.py 65536 2>&1 | grep "^B " >> $L; done
cat $L
