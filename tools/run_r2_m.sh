#!/bin/bash
# 2 GPUs: NCCL test of both shards, then the bench at N = 2 through torchrun
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2m.log; : > $L
nvidia-smi -L >> $L
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q -x 2>&1 | tail -15 >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2m_bench2.json 2>> $L
tail -c 1200 gpurun_out/r2m_bench2.json >> $L
cat $L
