"""Write a standalone .cu of the NVRTC translation unit for offline nvcc/ptxas/cuobjdump inspection."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sysbio_modeling_b200 import model_files
from sysbio_modeling_b200.symbolic import symbolic_model_from_file, emit_cuda
name, W, C, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
extra = sys.argv[5] if len(sys.argv) > 5 else ''
sm = symbolic_model_from_file(model_files.path(name), name=name)
cs = emit_cuda(sm)
H = os.path.join(ROOT, 'sysbio_modeling_b200', 'csrc')
src = open(os.path.join(H, 'sbm_kernel_args.h')).read() + open(os.path.join(H, 'sbm_device_math.h')).read() + \
    cs.source + open(os.path.join(H, 'kernels', 'sbm_ode_kernels.cuh')).read() + \
    open(os.path.join(H, 'kernels', 'sbm_dp5_ct.cuh')).read() + '\nSBM_DEFINE_DP5_CT(sbm_dp5_ct, %d, %d, %d)\n' % (W, C, int(os.environ.get("XW", "1"))) + extra
open(out, 'w').write(src)
