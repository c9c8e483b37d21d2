#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r16_tests.log; tail -4 gpurun_out/r16_tests.log
L=gpurun_out/r16_prof.log; : > $L
echo "== default (head 4)" >> $L; python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== head 1" >> $L; SBM_CT_HEAD_SHARES=1 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== head 2" >> $L; SBM_CT_HEAD_SHARES=2 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== head 6" >> $L; SBM_CT_HEAD_SHARES=6 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== head 4 shares 2" >> $L; SBM_CT_SHARES=2 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
cat $L
