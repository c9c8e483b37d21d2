from .base_project import Project
from .loss_functions import SquareLossFunction, LogSquareLossFunction, NormalizedSquareLossFunction

__all__ = ['Project', 'SquareLossFunction', 'LogSquareLossFunction', 'NormalizedSquareLossFunction']
