#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r5_variants.log; : > $L
echo "== C=3 W=8" >> $L; python scratch/prof_ct.py 65536 0,0,2 >> $L 2>&1
echo "== C=2 W=11" >> $L; python scratch/prof_ct.py 65536 0,2,2 >> $L 2>&1
echo "== C=1 W=22 shares=22" >> $L; python scratch/prof_ct.py 65536 0,1,2 >> $L 2>&1
echo "== C=1 W=22 shares=8" >> $L; SBM_CT_SHARES=8 python scratch/prof_ct.py 65536 0,1,2 >> $L 2>&1
echo "== C=2 W=11 shares=4" >> $L; SBM_CT_SHARES=4 python scratch/prof_ct.py 65536 0,2,2 >> $L 2>&1
grep -v "^{" $L
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -k "cta_transposed or c4_mapk" 2>&1 | tail -3
