#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2ah.log; : > $L
run() { echo "== $1 | $2" >> $L; SBM_KERNEL_DEFINES="$1" timeout 600 python tools/prof_ct.py ${2:-65536} 0,0,0 1e-10 2 2>&1 | grep -E "^B |rror|SBM_CT_PROF" >> $L; }
run "" 65536
run "-DSBM_CT_MODEL_PW=0" 65536
run "-DSBM_CT_MODEL_PW=3" 65536
run "-DSBM_CT_MODEL_PW=8" 65536
run "-DSBM_CT_MODEL_PW=7" 65536
cat $L
