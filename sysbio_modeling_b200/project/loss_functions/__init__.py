from .squared_loss import SquareLossFunction, LogSquareLossFunction, NormalizedSquareLossFunction

__all__ = ['SquareLossFunction', 'LogSquareLossFunction', 'NormalizedSquareLossFunction']
