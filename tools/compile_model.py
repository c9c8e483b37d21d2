import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sysbio_modeling_b200 import model_files, engine
from sysbio_modeling_b200.symbolic import symbolic_model_from_file
name = sys.argv[1] if len(sys.argv) > 1 else 'mapk_kholodenko'
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sm = symbolic_model_from_file(model_files.path(name), name=name)
cm = engine.CompiledModel(sm, cache_dir=None, cols_per_lane=cols, sens_layout=2 if cols else 0)
log = cm.log
for blk in log.split('ptxas info    : Compiling entry function')[1:]:
    lines = blk.strip().split('\n')
    print(lines[0][:40], '|', ' '.join(l.strip().replace('ptxas info    :','').replace('ptxas         .','') for l in lines[1:] if 'registers' in l or 'spill' in l))
i = cm.info()
print({k: i[k] for k in ('ct_warps', 'ct_cols', 'ct_dyn_smem_bytes', 'cubin_bytes')})
if len(sys.argv) > 3:
    open(sys.argv[3], 'w').write(cm.cuda.source)
