#!/bin/bash
# side measurements of the final kernel: looser tolerances, the legacy warp-group layout, small batches
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/check2.log; : > $L
echo "== tol 1e-8 (32768 sets)" >> $L; python tools/prof_ct.py 32768 0,0,0 1e-8 2>&1 | grep -E "^B " >> $L
echo "== tol 1e-6 (32768 sets)" >> $L; python tools/prof_ct.py 32768 0,0,0 1e-6 2>&1 | grep -E "^B " >> $L
echo "== tol 1e-10, 4736 sets (one wave)" >> $L; python tools/prof_ct.py 4736 2>&1 | grep -E "^B " >> $L
echo "== tol 1e-10, 262144 sets" >> $L; python tools/prof_ct.py 262144 2>&1 | grep -E "^B " >> $L
echo "== legacy warp-group layout (16384 sets)" >> $L; python tools/prof_ct.py 16384 0,0,1 2>&1 | grep -E "^B " >> $L
cat $L
