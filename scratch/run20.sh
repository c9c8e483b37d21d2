#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r23_tests.log; tail -4 gpurun_out/r23_tests.log
timeout 600 python scratch/bench_configs.py > gpurun_out/r23_configs.json 2> gpurun_out/r23_configs.err; grep "solves_per_s\|C[23]_" gpurun_out/r23_configs.json; tail -2 gpurun_out/r23_configs.err
