#!/bin/bash
# late round-2 captures: sbm_rodas4_sens, the row-parallel loss kernels, sbm_dp5_sens on the EGFR-scale model, and the launch list
# of the default bench (configs included).  Every command runs once WITHOUT ncu first.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python tools/prof_extra.py c3jac > gpurun_out/r02b_c3jac.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sbm_rodas4_sens -s 1 -c 1 -o gpurun_out/r02_rodas4_sens python tools/prof_extra.py c3jac > gpurun_out/r02b_ncu_c3jac.log 2>&1
python tools/prof_extra.py c2wide > gpurun_out/r02b_c2wide.log 2>&1 && ncu --set full --clock-control none -k regex:sbm_loss_wide_jac -s 1 -c 1 -o gpurun_out/r02_loss_wide_jac python tools/prof_extra.py c2wide > gpurun_out/r02b_ncu_c2wide.log 2>&1
python tools/prof_extra.py egfr > gpurun_out/r02b_egfr.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sbm_dp5_sens -s 1 -c 1 -o gpurun_out/r02_dp5_sens_egfr python tools/prof_extra.py egfr > gpurun_out/r02b_ncu_egfr.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r02b_plain.json 2> gpurun_out/r02b_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches.csv $CMD > gpurun_out/r02b_ncu_launch.log 2>&1
ls -la gpurun_out/r02_rodas4_sens* gpurun_out/r02_loss_wide* gpurun_out/r02_dp5_sens* gpurun_out/r02b_*; cat gpurun_out/r02b_c3jac.log gpurun_out/r02b_c2wide.log gpurun_out/r02b_egfr.log
