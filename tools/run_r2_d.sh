#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2d.log; : > $L
timeout 60 tools/tcp_probe >> $L 2>&1
python -m pytest tests/test_gpu_epilogue.py -m gpu -q -x 2>&1 | tail -15 >> $L
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 >> $L
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench.json 2>> $L
python -c "
import json; d=json.load(open('gpurun_out/r2d_bench.json')); print(d['value'], d['e2e']['value'], d['kernels'])" >> $L 2>&1
cat $L
