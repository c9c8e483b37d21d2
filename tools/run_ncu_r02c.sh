#!/bin/bash
# final round-2 captures of the two kernels that changed after tools/run_ncu_r02.sh: sbm_rodas4_plain (C3) and sbm_dp5_push (C5)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python tools/prof_configs.py c3 > gpurun_out/r02c_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sbm_rodas4_plain -s 2 -c 1 -o gpurun_out/r02_rodas4_plain_final python tools/prof_configs.py c3 > gpurun_out/r02c_ncu_c3.log 2>&1
python tools/prof_configs.py c5 > gpurun_out/r02c_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sbm_dp5_push -s 2 -c 1 -o gpurun_out/r02_dp5_push python tools/prof_configs.py c5 > gpurun_out/r02c_ncu_c5.log 2>&1
ls -la gpurun_out/r02_rodas4_plain_final* gpurun_out/r02_dp5_push*; cat gpurun_out/r02c_c3.log gpurun_out/r02c_c5.log
