#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r21_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r21_smoke.log
timeout 300 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "cta_transposed" 2>&1 | tail -5 > gpurun_out/r21_t1.log; echo "t1 rc=$?"; tail -3 gpurun_out/r21_t1.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r21_tests.log; tail -4 gpurun_out/r21_tests.log
L=gpurun_out/r21_prof.log; : > $L
echo "== TMEM" >> $L; timeout 300 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
echo "== smem" >> $L; SBM_CT_TMEM=0 timeout 300 python scratch/prof_ct.py 65536 2>&1 | grep "^B " >> $L
cat $L
