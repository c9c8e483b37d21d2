#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2q.log; : > $L
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 >> $L
python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "== default bench" >> $L
( time python bench.py ) > gpurun_out/r2q_bench.json 2>> $L
cat gpurun_out/r2q_bench.json >> $L
echo "== reference arm" >> $L
( time python bench.py --impl reference --steps 2 --warmup 1 ) >> $L 2>&1
echo "== ensemble" >> $L
python tools/bench_ensemble.py 131072 5 >> $L 2>&1
cat $L
