#!/bin/bash
# first GPU run of the CTA-transposed kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2_tests.log
tail -5 gpurun_out/r2_tests.log
timeout 300 python scratch/debug_loss.py > gpurun_out/r2_debug_loss.log 2>&1
for shape in 0,0,1 0,0,2; do
  timeout 600 python bench.py --batch 16384 --steps 2 --warmup 1 --no-cpu-baseline --shape $shape > gpurun_out/r2_bench_$shape.json 2> gpurun_out/r2_bench_$shape.err
  tail -c 1500 gpurun_out/r2_bench_$shape.json
done
