"""Host helpers of the project layer.

``OrderedHashDict`` keeps the scale-factor groups: keys are measurement names or frozensets of
names, and a name that lives inside a frozenset key resolves to that key's value (reference
``project/utils.py:98-158``, semantics pinned by reference ``tests/test_Project_Utils.py``).
The simulation->measurement mapping functions of reference ``utils.py:10-89`` do not exist here as
Python: they are the observable bit-masks evaluated inside the integrator kernels
(``csrc/kernels/sbm_ode_kernels.cuh: sbm_emit_rows``).
"""
from collections import OrderedDict

import numpy as np


class OrderedHashDict(OrderedDict):
    """Ordered dict keyed by ``str`` or ``frozenset[str]``; member strings alias their group."""

    def _group_of(self, name):
        for k in OrderedDict.keys(self):
            if isinstance(k, frozenset) and name in k:
                return k
        return None

    def __getitem__(self, key):
        if OrderedDict.__contains__(self, key):
            return OrderedDict.__getitem__(self, key)
        g = self._group_of(key)
        if g is None:
            raise KeyError('%s not in dictionary or in any of the groups in the dictionary' % (key,))
        return OrderedDict.__getitem__(self, g)

    def __contains__(self, key):
        return OrderedDict.__contains__(self, key) or self._group_of(key) is not None

    def __setitem__(self, key, value):
        if type(key) is not str and type(key) is not frozenset:
            raise TypeError('Keys can only be strings, or frozen sets of strings')
        if type(key) is frozenset:
            if not OrderedDict.__contains__(self, key):
                for member in key:
                    if type(member) is not str:
                        raise TypeError('Every element within the frozenset has to be a string')
                    if member in self:
                        raise KeyError('%s already in dict in a hashgroup' % (key,))
            return OrderedDict.__setitem__(self, key, value)
        if not OrderedDict.__contains__(self, key) and self._group_of(key) is not None:
            raise KeyError('%s already in dict in a hashgroup' % key)
        return OrderedDict.__setitem__(self, key, value)


def exp_param_transform(project_param_vector):
    """Parameters are optimised in log space: the model sees exp(theta) (reference utils.py:161-182)."""
    return np.exp(project_param_vector)


def exp_param_transform_derivative(project_param_vector):
    """d exp(theta)/d theta (reference utils.py:185-205); applied in-kernel as a multiply by p."""
    return np.exp(project_param_vector)
