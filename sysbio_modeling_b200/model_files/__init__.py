"""Model files (``#*!`` format of the reference's symbolic front end) for the BASELINE configs."""
import importlib
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def path(name):
    p = os.path.join(_HERE, name + '.py')
    if not os.path.exists(p):
        raise KeyError('no model file %s' % name)
    return p


def module(name):
    return importlib.import_module('sysbio_modeling_b200.model_files.' + name)
