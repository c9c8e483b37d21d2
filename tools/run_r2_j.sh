#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2j.log; : > $L
echo "== LAG=1 tables in smem 65536" >> $L; timeout 300 python tools/prof_ct.py 65536 2>&1 | grep -E "^B |rror" >> $L
echo "== PROF" >> $L; SBM_KERNEL_DEFINES="-DSBM_CT_PROF=0" timeout 300 python tools/prof_ct.py 16384 0,0,0 1e-10 1 2>&1 | grep -E "^B |rror|SBM_CT_PROF" | tail -5 >> $L
timeout 900 python -m pytest tests/test_gpu_epilogue.py tests/test_gpu_bench_parity.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4 >> $L
echo "== bench" >> $L
timeout 900 python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline 2>>$L | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'],'serial',d['e2e']['serial_single_call'], d['e2e']['matches_device_path']); print(d['kernels']); print(d['roofline']['frac'], d['roofline']['frac_emitter_count'])" >> $L 2>&1
cat $L
