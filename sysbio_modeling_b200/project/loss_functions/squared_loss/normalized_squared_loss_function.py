"""Normalized square loss: sigma_i <- sigma_i * Y_i
(reference ``squared_loss/normalized_squared_loss_function.py:6-46``)."""
from .squared_loss_function import SquareLossFunction


class NormalizedSquareLossFunction(SquareLossFunction):
    kind = 'normalized'
