"""Optimal linear scale factor B = sum(s d/sigma^2)/sum(s^2/sigma^2)
(reference ``squared_loss/linear_scale_factor.py:21-88``); evaluated on the GPU."""
import numpy as np

from ..abstract_scale_factor import ScaleFactorABC


class LinearScaleFactor(ScaleFactorABC):
    kind = 'linear'

    def __init__(self, log_prior=None, log_prior_sigma=None):
        super(LinearScaleFactor, self).__init__(log_prior, log_prior_sigma)
        self._sf = 1.0

    def calc_scale_factor_entropy(self, sim_dot_sim, sim_dot_exp, temperature=1.0):
        """log of the SloppyCell scale-factor integral (reference :63-80, integrand :13-18):

            int exp(-a/(2T) (B e^u - B)^2 - (u + log B - mu)^2 / (2 sg^2)) du,   a = sum(s^2/sigma^2)

        evaluated by composite 16-point Gauss-Legendre panels over a window that covers both the
        likelihood peak (u = 0) and the prior peak (u = mu - log B), panel width tied to the narrower
        of the two.  The reference calls QUADPACK; this is host code for the ensemble sampler, not part
        of the residual hot path."""
        if self._sf is None:
            return 0
        mu, sg = self.log_prior, self.log_prior_sigma
        if mu is None or sg is None:
            raise TypeError('scale-factor entropy needs a log prior (as in the reference)')
        B = sim_dot_exp / sim_dot_sim
        lB = np.log(B)
        w_like = np.sqrt(temperature / (sim_dot_sim * B * B))
        width = max(min(w_like, sg), 1e-9)
        lo = min(0.0, mu - lB) - 12.0 * max(w_like, sg)
        hi = max(0.0, mu - lB) + 12.0 * max(w_like, sg)
        n_panels = int(min(max(np.ceil((hi - lo) / width), 8), 20000))
        edges = np.linspace(lo, hi, n_panels + 1)
        x, w = np.polynomial.legendre.leggauss(16)
        half = 0.5 * (edges[1:] - edges[:-1])
        mid = 0.5 * (edges[1:] + edges[:-1])
        u = mid[:, None] + half[:, None] * x[None, :]
        f = np.exp(-sim_dot_sim / (2 * temperature) * (B * np.exp(u) - B) ** 2
                   - (u + lB - mu) ** 2 / (2 * sg ** 2))
        return np.log(np.sum(half[:, None] * w[None, :] * f))
