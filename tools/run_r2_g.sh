#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2g.log; : > $L
timeout 900 python -m pytest tests/test_gpu_determinism.py -m gpu -q -x 2>&1 | tail -12 >> $L
echo "== default 65536 (J slots shared)" >> $L; timeout 300 python tools/prof_ct.py 65536 2>&1 | grep -E "^B |rror" >> $L
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 >> $L
echo "== bench e2e" >> $L
timeout 600 python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline 2>>$L | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['matches_device_path'], d['kernels']['sbm_loss_tile_kernel_ms'])" >> $L 2>&1
cat $L
