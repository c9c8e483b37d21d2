"""Square loss (reference ``squared_loss/squared_loss_function.py:9-106``), computed on the GPU."""
import numpy as np
import pandas as pd

from ..abstract_loss_function import LossFunctionWithScaleFactors, DifferentiableLossFunctionABC
from .linear_scale_factor import LinearScaleFactor


class SquareLossFunction(LossFunctionWithScaleFactors, DifferentiableLossFunctionABC):
    """C(theta) = 0.5 * sum(((B X_i - Y_i)/sigma_i)^2) with optional optimal scale factors B."""
    kind = 'square'

    def __init__(self, sf_groups=None, sf_type=LinearScaleFactor):
        super(SquareLossFunction, self).__init__(sf_groups, sf_type)

    def evaluate(self, simulations, experiment_measures):
        res = self.residuals(simulations, experiment_measures)
        return 0.5 * np.sum(np.asarray(res) ** 2)

    def residuals(self, simulations, experiment_measures):
        """Reference :27-42.  Returns a Series in the row order of ``simulations``; all-inf ndarray when
        the simulations contain NaN."""
        out = self._run_frames(simulations, experiment_measures)
        if np.isinf(out['resid']).all() and np.isnan(np.asarray(simulations['mean'].values, dtype=float)).any():
            return out['resid']
        return pd.Series(out['resid'], index=simulations.index)

    def update_scale_factors_gradient(self, simulations, experiment_measures, simulations_jacobian):
        """Reference :82-106."""
        if len(self._scale_factors) == 0:
            return
        self._run_frames(simulations, experiment_measures, simulations_jacobian)

    def jacobian(self, simulations, experiment_measures, simulations_jacobian):
        """Reference :44-80: B*J + sim*dB per scaled measure; unchanged without scale factors."""
        out = self._run_frames(simulations, experiment_measures, simulations_jacobian)
        return pd.DataFrame(out['jac'], index=simulations_jacobian.index, columns=simulations_jacobian.columns)
