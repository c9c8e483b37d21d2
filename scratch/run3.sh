#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r4_tests.log
tail -6 gpurun_out/r4_tests.log
python scratch/prof_ct.py 16384 > gpurun_out/r4_prof.log 2>&1
python scratch/prof_ct.py 65536 >> gpurun_out/r4_prof.log 2>&1
cat gpurun_out/r4_prof.log
