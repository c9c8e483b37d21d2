#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2n.log; : > $L
timeout 900 python -m pytest tests/test_ensembles.py -q -x 2>&1 | tail -12 >> $L
timeout 600 python tools/bench_ensemble.py 131072 5 >> $L 2>&1
cat $L
