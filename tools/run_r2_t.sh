#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2t.log; : > $L
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 >> $L
echo "== bench (full configs)" >> $L
python bench.py --steps 2 --warmup 1 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], json.dumps(d['configs']['c5'],indent=1), json.dumps(d['configs']['c3']))" >> $L 2>&1
cat $L
