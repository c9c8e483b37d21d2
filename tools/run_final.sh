#!/bin/bash
# final artefacts of a round: launch list + full ncu capture of the default bench command (batch 16384), the default
# bench line (65,536 sets), the other configs
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
CMD="python bench.py --batch 16384 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/final_plain.json 2> gpurun_out/final_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/final_ncu_launch.log 2>&1
$CMD > gpurun_out/final_plain2.json 2> gpurun_out/final_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:sbm_dp5 -s 3 -c 1 -o gpurun_out/prof_final $CMD > gpurun_out/final_ncu_full.log 2>&1
tail -c 300 gpurun_out/final_plain.json; tail -2 gpurun_out/final_ncu_full.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cat gpurun_out/final_bench.json
python tools/bench_configs.py > gpurun_out/final_configs.json 2> gpurun_out/final_configs.err; cat gpurun_out/final_configs.json
SBM_KERNEL_DEFINES=-DSBM_CT_PROF=0 python tools/prof_ct.py 65536 2>&1 | grep -E "SBM_CT_PROF|^B " | tail -6 > gpurun_out/final_intervals.txt; cat gpurun_out/final_intervals.txt
