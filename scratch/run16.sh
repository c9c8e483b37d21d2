#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r18_tests.log; tail -4 gpurun_out/r18_tests.log
L=gpurun_out/r18_prof.log; : > $L
echo "== G as LDS.128" >> $L; SBM_CT_HEAD_SHARES=2 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
cat $L
