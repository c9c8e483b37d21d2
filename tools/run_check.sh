#!/bin/bash
# GPU parity suite + timing of the default C4 configuration (and optional variants given as "ENV=.. ENV=.." lines in $1)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/check.log; : > $L
[ "$2" = "notest" ] || python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $L
echo "== default" >> $L; python tools/prof_ct.py 65536 2>&1 | grep -E "^B |ct_warps|rror|SBM_CT_PROF" >> $L
if [ -n "$1" ]; then
  while IFS= read -r line; do
    [ -z "$line" ] && continue
    echo "== $line" >> $L; env "$line" python tools/prof_ct.py 65536 2>&1 | grep -E "^B |ct_warps|rror|SBM_CT_PROF" >> $L
  done < "$1"
fi
cat $L
