#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
SBM_KERNEL_DEFINES="-DSBM_CT_PROF=1" python scratch/prof_ct.py 16384 > gpurun_out/r3_prof.log 2>&1
python scratch/prof_ct.py 16384 >> gpurun_out/r3_prof.log 2>&1
python scratch/prof_ct.py 65536 >> gpurun_out/r3_prof.log 2>&1
cat gpurun_out/r3_prof.log
