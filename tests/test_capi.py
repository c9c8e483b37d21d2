"""CPU: the C-ABI library builds, loads, and exports every symbol include/sbm_b200.h declares; compute entry
points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'sbm_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sbm_[a-z0-9_]+)\s*\(', text)))


def test_exports_match_header():
    from sysbio_modeling_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), 'library does not export %s' % n
    assert sorted(_lib.EXPORTS) == names
    assert L.sbm_version() >= 100


def test_nvrtc_compile_without_gpu_and_cache(tmp_path):
    """NVRTC cross-compiles the model + integrator kernels for sm_100a with no device present."""
    from sysbio_modeling_b200 import engine, model_files
    from sysbio_modeling_b200.symbolic import symbolic_model_from_file
    sm = symbolic_model_from_file(model_files.path('michaelis_menten'))
    cm = engine.CompiledModel(sm, cache_dir=str(tmp_path))
    assert 'sbm_dp5_sens' in cm.log and 'sbm_rodas4_plain' in cm.log
    assert len(os.listdir(tmp_path)) >= 1
    cm2 = engine.CompiledModel(sm, cache_dir=str(tmp_path))       # second time comes from the cache
    from sysbio_modeling_b200 import _lib
    if _lib.device_count() == 0:
        info = _lib.ModelInfo()
        assert _lib.lib().sbm_model_info(cm2.handle, ctypes.byref(info)) == 0
        assert info.from_cache == 1 and info.regs_dp5_sens == -1


def test_bad_model_source_reports_nvrtc_log():
    from sysbio_modeling_b200 import _lib
    d = _lib.ModelDesc(1, 1, 1, 0, 0, 0, -1, None, None)
    h = ctypes.c_void_p()
    rc = _lib.lib().sbm_model_compile(b'this is not CUDA', ctypes.byref(d), ctypes.byref(h))
    assert rc == -2
    assert b'error' in _lib.lib().sbm_last_error()


def test_no_cpu_fallback():
    from sysbio_modeling_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip('GPU present')
    from fixtures import project_1state
    proj, _ = project_1state()
    with pytest.raises(_lib.SbmError):
        proj.residuals(np.zeros(3))
    with pytest.raises(_lib.SbmError):
        _lib.fp64_peak()
