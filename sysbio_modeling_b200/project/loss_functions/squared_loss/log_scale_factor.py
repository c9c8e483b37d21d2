"""Geometric-mean scale factor of the log loss (reference ``squared_loss/log_scale_factor.py:9-82``);
evaluated on the GPU."""
from ..abstract_scale_factor import ScaleFactorABC


class LogScaleFactor(ScaleFactorABC):
    kind = 'log'

    def __init__(self, log_prior=None, log_prior_sigma=None):
        super(LogScaleFactor, self).__init__(log_prior, log_prior_sigma)
        self._sf = 0
