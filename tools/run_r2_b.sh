#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_gpu_bench_parity.py -m gpu -q -s 2>&1 | tail -25 > gpurun_out/r2b_parity.log
bash tools/run_whatif.sh > /dev/null 2>&1
cat gpurun_out/r2b_parity.log gpurun_out/whatif.log
