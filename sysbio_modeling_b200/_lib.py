"""ctypes binding of libsbm_b200.so (C ABI in include/sbm_b200.h).

The library is built in-tree by ``csrc/build.py``.  There is no fallback: if the shared object is
missing, or no CUDA device is visible when a compute entry point is called, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsbm_b200.so')
CACHE_DIR = os.path.join(_HERE, '_cache')

SBM_SOLVER_DP5, SBM_SOLVER_RODAS4 = 0, 1
SBM_LOSS_SQUARE, SBM_LOSS_LOG, SBM_LOSS_NORMALIZED = 0, 1, 2
SBM_FLAG_PRIOR_JAC_UNIT = 1
SBM_FLAG_LOSS_GENERIC = 2
SBM_ROW_MEASUREMENT, SBM_ROW_PARAM_PRIOR, SBM_ROW_SF_PRIOR, SBM_ROW_PLAIN = 0, 1, 2, 3
SBM_N_COUNTERS = 8
STATUS_NAMES = {0: 'ok', 1: 'max_steps', 2: 'step_underflow', 3: 'non_finite'}
COUNTER_NAMES = ('steps', 'rejects', 'rhs_evals', 'lu', 'trajectories')


class SbmError(RuntimeError):
    pass


class ModelDesc(ctypes.Structure):
    _fields_ = [('n_vars', ctypes.c_int), ('n_params', ctypes.c_int), ('n_free', ctypes.c_int),
                ('group_lanes', ctypes.c_int), ('cols_per_lane', ctypes.c_int), ('block_threads', ctypes.c_int),
                ('sens_layout', ctypes.c_int), ('arch', ctypes.c_char_p), ('cache_dir', ctypes.c_char_p)]


class ModelInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        'n_vars', 'n_params', 'n_free', 'group_lanes', 'cols_per_lane', 'block_threads', 'stage_smem',
        'dyn_smem_bytes',
        'regs_dp5_plain', 'regs_dp5_sens', 'regs_rodas4',
        'local_dp5_plain', 'local_dp5_sens', 'local_rodas4',
        'smem_dp5_plain', 'smem_dp5_sens', 'smem_rodas4', 'cubin_bytes', 'from_cache',
        'sens_layout', 'ct_warps', 'ct_cols', 'ct_cols_state', 'ct_dyn_smem_bytes', 'ct_ctas_per_sm',
        'regs_dp5_ct', 'local_dp5_ct', 'smem_dp5_ct')]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_IP = ctypes.POINTER(ctypes.c_int)
_DP = ctypes.POINTER(ctypes.c_double)
_UP = ctypes.POINTER(ctypes.c_ulonglong)


class PlanDesc(ctypes.Structure):
    _fields_ = [('E', ctypes.c_int), ('n_rows', ctypes.c_int), ('n_proj', ctypes.c_int),
                ('n_groups', ctypes.c_int), ('n_tout', ctypes.c_int), ('n_orow', ctypes.c_int),
                ('jac_atomic', ctypes.c_int),
                ('pmap', _IP), ('pfix', _DP), ('y0', _DP), ('tout_off', _IP), ('tout', _DP),
                ('orow_off', _IP), ('row_sim', _IP), ('row_jac', _IP), ('row_mask', _UP),
                ('data', _DP), ('sigma', _DP), ('row_group', _IP), ('row_kind', _IP), ('row_aux', _IP),
                ('grp_off', _IP), ('grp_rows', _IP)]


class SolverOpts(ctypes.Structure):
    _fields_ = [('solver', ctypes.c_int), ('max_steps', ctypes.c_int), ('jac_raw', ctypes.c_int),
                ('dense_output', ctypes.c_int), ('rtol', ctypes.c_double),
                ('atol', ctypes.c_double), ('h0', ctypes.c_double)]


EXPORTS = ['sbm_model_compile', 'sbm_model_info', 'sbm_model_log', 'sbm_model_destroy', 'sbm_plan_create',
           'sbm_plan_destroy', 'sbm_solve', 'sbm_loss', 'sbm_loss_sums', 'sbm_loss_finish', 'sbm_normal_equations',
           'sbm_residuals_host', 'sbm_jacobian_host', 'sbm_jacobian_host_begin', 'sbm_host_end', 'sbm_lm_step', 'sbm_mc_propose', 'sbm_mc_accept',
           'sbm_fp64_peak', 'sbm_device_count', 'sbm_last_error', 'sbm_version']

_lib = None


def lib():
    """Load the native library (once).  Raises SbmError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SbmError('libsbm_b200.so is not built (run python sysbio_modeling_b200/csrc/build.py or '
                       '__graft_entry__.build()); this engine has no CPU fallback')
    L = ctypes.CDLL(LIB_PATH)
    vp, ll, ci, cd = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_double
    L.sbm_last_error.restype = ctypes.c_char_p
    L.sbm_model_log.restype = ctypes.c_char_p
    L.sbm_model_log.argtypes = [vp]
    L.sbm_model_compile.argtypes = [ctypes.c_char_p, ctypes.POINTER(ModelDesc), ctypes.POINTER(vp)]
    L.sbm_model_info.argtypes = [vp, ctypes.POINTER(ModelInfo)]
    L.sbm_model_destroy.argtypes = [vp]
    L.sbm_model_destroy.restype = None
    L.sbm_plan_create.argtypes = [vp, ctypes.POINTER(PlanDesc), ctypes.POINTER(vp)]
    L.sbm_plan_destroy.argtypes = [vp]
    L.sbm_plan_destroy.restype = None
    L.sbm_solve.argtypes = [vp, ctypes.POINTER(SolverOpts), ll, vp, ll, ll, vp, vp, vp, vp, vp]
    L.sbm_loss.argtypes = [vp, ci, ci, ll, vp, ll, ll, vp, vp, vp, vp, vp, vp, vp, vp]
    L.sbm_loss_sums.argtypes = [vp, ci, ci, ll, vp, ll, ll, vp, vp, vp, vp]
    L.sbm_loss_finish.argtypes = [vp, ci, ci, ll, vp, ll, ll, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.sbm_normal_equations.argtypes = [ll, ci, ci, vp, vp, vp, vp]
    L.sbm_residuals_host.argtypes = [vp, ctypes.POINTER(SolverOpts), ci, ll, vp, vp, vp, vp, vp, vp, vp, vp]
    L.sbm_jacobian_host.argtypes = [vp, ctypes.POINTER(SolverOpts), ci, ci, ll, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.sbm_jacobian_host_begin.argtypes = [vp, ctypes.POINTER(SolverOpts), ci, ci, ll, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ci]
    L.sbm_host_end.argtypes = [vp, ci]
    L.sbm_lm_step.argtypes = [ll, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    ull, cu = ctypes.c_ulonglong, ctypes.c_uint
    L.sbm_mc_propose.argtypes = [ll, ci, vp, vp, vp, vp, ull, cu, vp]
    L.sbm_mc_accept.argtypes = [vp, ll, ci, cd, vp, vp, vp, vp, vp, vp, vp, vp, ull, cu, vp, vp, vp, vp]
    L.sbm_fp64_peak.argtypes = [ci, _DP, _DP]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise SbmError('libsbm_b200 error %d: %s' % (rc, lib().sbm_last_error().decode(errors='replace')))


def device_count():
    return lib().sbm_device_count()


def require_gpu():
    if device_count() <= 0:
        raise SbmError('no CUDA device visible: the sysbio_modeling_b200 CUDA backend has no CPU fallback')


def fp64_peak(iters=8192):
    tf, ms = ctypes.c_double(), ctypes.c_double()
    check(lib().sbm_fp64_peak(iters, ctypes.byref(tf), ctypes.byref(ms)))
    return tf.value, ms.value
