#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
CMD="python bench.py --batch 16384 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r11_plain.json 2> gpurun_out/r11_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r11_launches.csv $CMD > gpurun_out/r11_ncu_launch.log 2>&1
tail -c 400 gpurun_out/r11_plain.json
$CMD > gpurun_out/r11_plain2.json 2> gpurun_out/r11_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:sbm_dp5 -s 3 -c 1 -o gpurun_out/prof_r1_ct_v3 $CMD > gpurun_out/r11_ncu_full.log 2>&1
tail -3 gpurun_out/r11_ncu_full.log
