#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2l.log; : > $L
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 >> $L
python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
cat $L
