#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2ak.log; : > $L
timeout 1200 python -m pytest tests/test_gpu_stiff_sens.py tests/test_gpu_dense_output.py tests/test_gpu_configs.py tests/test_gpu_bench_parity.py tests/test_gpu_full_size.py -m gpu -q -x 2>&1 | tail -5 >> $L
python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], json.dumps(d['configs']['c3'],indent=1))" >> $L 2>&1
cat $L
