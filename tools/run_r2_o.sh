#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2o.log; : > $L
for nb in 1 2; do for w in 6 8 12; do SBM_LOSS_NBUF=$nb SBM_LOSS_WARPS=$w timeout 200 python tools/bench_loss.py 2>&1 | grep -E "loss ms|rror" >> $L; done; done
timeout 200 python tools/bench_loss.py 2>&1 | grep -E "ms|rror" >> $L
timeout 600 python -m pytest tests/test_gpu_epilogue.py tests/test_ensembles.py -q -x 2>&1 | tail -3 >> $L
timeout 600 python tools/bench_ensemble.py 131072 5 >> $L 2>&1
cat $L
