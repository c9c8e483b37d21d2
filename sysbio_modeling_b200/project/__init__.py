from .base_project import Project
from .loss_functions import SquareLossFunction, LogSquareLossFunction, NormalizedSquareLossFunction
from . import ensembles

__all__ = ['Project', 'SquareLossFunction', 'LogSquareLossFunction', 'NormalizedSquareLossFunction', 'ensembles']
