#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r9_tests.log; tail -4 gpurun_out/r9_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err; tail -c 3000 gpurun_out/r9_bench.json; tail -3 gpurun_out/r9_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r9_bench_ref.json 2> gpurun_out/r9_bench_ref.err; tail -c 1200 gpurun_out/r9_bench_ref.json; tail -3 gpurun_out/r9_bench_ref.err
