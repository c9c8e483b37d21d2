#!/bin/bash
# probe: one column per warp (22 column warps + model warps) against the default 11 x 2 layout
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/probe.log; : > $L
run() { echo "== $1" >> $L; shift; env "$@" 2>&1 | grep -E "^B |ct_warps|Error|error" >> $L; }
run "default"                 X=1 python tools/prof_ct.py 65536
run "c1 xw2 shares2"          SBM_CT_XW=2 SBM_CT_SHARES=2 python tools/prof_ct.py 65536 0,1,2
run "c1 xw1 shares1"          SBM_CT_XW=1 python tools/prof_ct.py 65536 0,1,2
run "c1 xw2 shares2 head4"    SBM_CT_XW=2 SBM_CT_SHARES=2 SBM_CT_HEAD_SHARES=4 python tools/prof_ct.py 65536 0,1,2
run "c1 xw3 shares3"          SBM_CT_XW=3 SBM_CT_SHARES=3 python tools/prof_ct.py 65536 0,1,2
cat $L
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
