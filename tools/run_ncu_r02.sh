#!/bin/bash
# round-2 profiles: launch list + full capture of the headline kernel at the bench configuration, full captures of the
# loss / normal-equation kernels and of the C3 / C5 integrators.  Every command runs once WITHOUT ncu first.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline"
$CMD > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sbm_dp5_ct -s 4 -c 1 -o gpurun_out/r02_dp5_ct $CMD > gpurun_out/r02_ncu_ct.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sbm_loss_tile -s 4 -c 1 -o gpurun_out/r02_loss_tile $CMD > gpurun_out/r02_ncu_loss.log 2>&1
ncu --set full --clock-control none -k regex:sbm_gram -s 4 -c 1 -o gpurun_out/r02_gram $CMD > gpurun_out/r02_ncu_gram.log 2>&1
python tools/prof_configs.py c3 > gpurun_out/r02_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sbm_rodas4 -s 2 -c 1 -o gpurun_out/r02_rodas4_plain python tools/prof_configs.py c3 > gpurun_out/r02_ncu_c3.log 2>&1
python tools/prof_configs.py c5 > gpurun_out/r02_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sbm_dp5_plain -s 2 -c 1 -o gpurun_out/r02_dp5_plain python tools/prof_configs.py c5 > gpurun_out/r02_ncu_c5.log 2>&1
ls -la gpurun_out/r02_*; tail -2 gpurun_out/r02_ncu_ct.log
