#!/bin/bash
# 8 GPUs: the bench through torchrun (no side configs), the reference arm launched the same way, the sharded ensemble
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2v.log; : > $L
nvidia-smi -L >> $L
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 3 --warmup 3 --no-configs > gpurun_out/r2v_bench$N.json 2>> $L
tail -c 3000 gpurun_out/r2v_bench$N.json >> $L
echo >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 tools/bench_ensemble.py 16384 5 >> $L 2>&1
cat $L | tail -c 6000
