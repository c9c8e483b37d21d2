#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2f.log; : > $L
echo "== ASYNC=1 65536" >> $L; SBM_CT_ASYNC=1 timeout 300 python tools/prof_ct.py 65536 2>&1 | grep -E "^B |rror" >> $L
echo "== ASYNC=1 NOMODEL 16384 (max 3000 steps)" >> $L; SBM_PROF_MAXSTEPS=3000 SBM_CT_ASYNC=1 SBM_KERNEL_DEFINES=-DSBM_WHATIF_ASYNC_NOMODEL timeout 300 python tools/prof_ct.py 16384 0,0,0 1e-10 1 2>&1 | grep -E "^B |rror" >> $L
echo "== sync default 16384 (max 3000 steps)" >> $L; SBM_PROF_MAXSTEPS=3000 timeout 300 python tools/prof_ct.py 16384 0,0,0 1e-10 1 2>&1 | grep -E "^B |rror" >> $L
echo "== bench quick" >> $L
timeout 900 python bench.py --steps 2 --warmup 3 --quick-configs > gpurun_out/r2f_bench.json 2>> $L
python - >> $L 2>&1 <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['matches_device_path'])
print('kernels',d['kernels'])
print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','frac_emitter_count')})
print('parity',d.get('parity'))
print('configs',json.dumps(d.get('configs'),indent=1))
print('cpu',d.get('cpu_baseline'))
PY
echo "== e2e serial" >> $L
SBM_HOST_CHUNKS=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline 2>>$L | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e serial',d['e2e']['value'])" >> $L 2>&1
cat $L
