#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; shift
CMD="python bench.py --batch 4736 --steps 1 --warmup 1 --no-cpu-baseline $@"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:sbm_dp5 -s 1 -c 1 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -c 600 gpurun_out/${TAG}_plain.json
tail -3 gpurun_out/${TAG}_ncu.log
