#!/bin/bash
# what-if builds of sbm_dp5_ct (wrong results, bounded steps): where a stage's time goes
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/whatif.log; : > $L
run() { echo "== $1" >> $L; SBM_PROF_MAXSTEPS=${3:-0} SBM_KERNEL_DEFINES="-DSBM_CT_PROF=0 $1" timeout 300 python tools/prof_ct.py ${2:-16384} 0,0,0 1e-10 1 2>&1 | grep -E "^B |rror|SBM_CT_PROF" >> $L; }
run ""
run "-DSBM_WHATIF_COLS_SKIP" 16384 3000
run "-DSBM_CT_PRIMAL_REPEAT=0" 16384 3000
run "-DSBM_JLOAD(q)=1e-30" 16384 3000
run "-DSBM_WHATIF_NO_G" 16384 3000
run "-DSBM_WHATIF_NO_TMEM" 16384 3000
run "-DSBM_JLOAD(q)=1e-30 -DSBM_WHATIF_NO_G -DSBM_WHATIF_NO_TMEM" 16384 3000
run "-DSBM_WHATIF_NO_COL_EMIT" 16384 3000
cat $L
