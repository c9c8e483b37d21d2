#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2r.log; : > $L
timeout 900 python -m pytest tests/test_gpu_epilogue.py -m gpu -q -x -s 2>&1 | tail -60 >> $L
echo "== bench quick" >> $L
python bench.py --steps 2 --warmup 1 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps(d['configs']['c2'],indent=1))" >> $L 2>&1
cat $L
