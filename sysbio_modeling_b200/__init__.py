"""sysbio_modeling_b200 - B200-native batched ODE-with-sensitivities engine behind the
OdeModel / Experiment / TimecourseMeasurement / Project API of FedericoV/SysBio_Modeling.

Package layout mirrors the reference's top-level packages for the hot path only:
``experiment``, ``measurement``, ``model``, ``project``, ``symbolic`` (+ ``csrc`` for the CUDA side).
"""
from .experiment import Experiment
from .measurement import TimecourseMeasurement
from .model import OdeModel
from .project import Project, SquareLossFunction, LogSquareLossFunction, NormalizedSquareLossFunction
from ._lib import SbmError

__all__ = ['Experiment', 'TimecourseMeasurement', 'OdeModel', 'Project', 'SquareLossFunction',
           'LogSquareLossFunction', 'NormalizedSquareLossFunction', 'SbmError']
__version__ = '0.1.0'
