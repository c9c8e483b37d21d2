"""Recipe for the reference arm of ``bench.py --impl reference``: copy the reference's OWN implementation of the path
- ``model/ode_model.py`` (``OdeModel.calc_jacobian`` / ``simulate``: SciPy LSODA on a numba-compiled RHS) and the base
class it imports, ``model/abstract_model.py`` - UNMODIFIED (byte for byte, sha256 recorded) from /root/reference into the
git-ignored ``baseline/_ref/model/``.  Nothing is patched: the only Python-2-ism of these two files is the implicit
relative import ``from abstract_model import ModelABC`` (ode_model.py:3), which is satisfied by putting that directory on
``sys.path``.  The rest of the reference (``project/``: Python 2 print statements, pandas-0.x calls) does not import on
Python 3 and the reference has no setup.py / pyproject.toml, so ``pip install --target baseline/_ref /root/reference`` is
not possible.

    python oracle/build_ref.py            # run by __graft_entry__.build() when /root/reference exists

``baseline/_ref`` travels to the GPU box with the repository snapshot; /root/reference itself is never read at run time.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference/model'
DST = os.path.join(ROOT, 'baseline', '_ref', 'model')
FILES = ('ode_model.py', 'abstract_model.py')


def build(verbose=True):
    if not os.path.isdir(SRC):
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        manifest[f] = hashlib.sha256(open(os.path.join(DST, f), 'rb').read()).hexdigest()
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as fh:
        json.dump({'source': SRC, 'sha256': manifest, 'modified': False}, fh, indent=1)
    if verbose:
        print('baseline/_ref/model: %s' % ', '.join('%s %s' % (k, v[:12]) for k, v in manifest.items()))
    return True


def load_reference_odemodel():
    """The reference's unmodified ``OdeModel`` class from baseline/_ref, or None when the copy is absent."""
    if not all(os.path.exists(os.path.join(DST, f)) for f in FILES):
        return None
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import importlib
    mod = importlib.import_module('ode_model')
    if os.path.dirname(os.path.abspath(mod.__file__)) != DST:
        return None
    return mod.OdeModel


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
