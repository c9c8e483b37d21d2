#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2w.log; : > $L
N=$(nvidia-smi -L | wc -l)
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 3 --warmup 3 ) > gpurun_out/r2w_bench$N.json 2>> $L
python - >> $L <<PY
import json
d=json.loads(open('gpurun_out/r2w_bench$N.json').read().strip().splitlines()[-1])
print(d['value'], d['n_gpus'], d['e2e']['value'], d.get('parity'), 'configs' in d, 'cpu_baseline' in d)
PY
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q -x 2>&1 | tail -3 >> $L
cat $L | tail -c 3000
