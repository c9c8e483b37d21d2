#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r13_tests.log; tail -4 gpurun_out/r13_tests.log
L=gpurun_out/r13_prof.log; : > $L
echo "== default (XW=1)" >> $L; python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== XW=0" >> $L; SBM_CT_XW=0 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== XW=1 shares=4" >> $L; SBM_CT_SHARES=4 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== XW=1 shares=2" >> $L; SBM_CT_SHARES=2 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
cat $L
