import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden():
    import json
    with open(os.path.join(ROOT, 'tests', 'golden', 'reference_golden.json')) as fh:
        return json.load(fh)


@pytest.fixture(scope='session', autouse=True)
def _native_built():
    """Build the native library once per session (a no-op when it is up to date)."""
    from sysbio_modeling_b200.csrc import build as native_build
    native_build.build(verbose=False)
