"""Log-space square loss (reference ``squared_loss/log_squared_loss_function.py:9-98``)."""
from .squared_loss_function import SquareLossFunction
from .log_scale_factor import LogScaleFactor


class LogSquareLossFunction(SquareLossFunction):
    """C(theta) = 0.5 * sum(((log(B X_i) - log(Y_i))/sigma_i)^2); geometric-mean scale factors."""
    kind = 'log'

    def __init__(self, sf_groups=None):
        super(LogSquareLossFunction, self).__init__(sf_groups, LogScaleFactor)
