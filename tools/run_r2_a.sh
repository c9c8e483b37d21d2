#!/bin/bash
# round 2, call A: new parity tests first (verbose), then the whole GPU suite, then the baseline bench line
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_gpu_bench_parity.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r2a_parity.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
cat gpurun_out/r2a_parity.log gpurun_out/r2a_tests.log; tail -c 1500 gpurun_out/r2a_bench.json
