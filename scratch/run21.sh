#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r24_tests.log; tail -6 gpurun_out/r24_tests.log
timeout 900 python bench.py > gpurun_out/r24_bench.json 2> gpurun_out/r24_bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r24_bench.json').read().strip().split('\n')[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline']['traffic'], d.get('cpu_baseline'), d['clocks'])
PY
tail -2 gpurun_out/r24_bench.err
