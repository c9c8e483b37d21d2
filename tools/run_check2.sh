#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/check2.log; : > $L
for cfg in "3 150" "3 100" "4 100" "5 100" "5 150" "4 200" "3 200"; do
  set -- $cfg
  echo "== head shares $1 colbase $2" >> $L
  SBM_CT_HEAD_SHARES=$1 SBM_CT_HEAD_COLBASE=$2 python tools/prof_ct.py 65536 2>&1 | grep -E "^B |rror" >> $L
done
cat $L
