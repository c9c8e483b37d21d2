"""Scale-factor value holders (reference ``project/loss_functions/abstract_scale_factor.py:6-40``).

The numbers are produced by the CUDA loss epilogue (``csrc/sbm_epilogue.cuh``); these objects only
carry the last evaluated value, its gradient and the optional log-prior, as the reference's do.
"""
import numpy as np


class ScaleFactorABC(object):
    def __init__(self, log_prior=None, log_prior_sigma=None):
        self._sf = None
        self._sf_gradient = None
        self.log_prior = log_prior
        self.log_prior_sigma = log_prior_sigma

    @property
    def sf(self):
        return self._sf

    @property
    def gradient(self):
        return None if self._sf_gradient is None else self._sf_gradient.copy()

    def calc_sf_prior_residual(self):
        """(log B - prior)/sigma (reference linear_scale_factor.py:55-61)."""
        if self.log_prior is None:
            return None
        return (np.log(self._sf) - self.log_prior) / self.log_prior_sigma

    def calc_sf_prior_gradient(self):
        """dB/B (reference linear_scale_factor.py:44-53)."""
        if self.log_prior is None:
            return None
        return self._sf_gradient / self._sf

    def __repr__(self):
        out = "SF value: %.5e\n" % self._sf
        if self.log_prior is not None:
            out += "SF log prior: %.4f\n" % self.log_prior
            out += "SF log prior sigma: %.4f\n" % self.log_prior_sigma
        return out
