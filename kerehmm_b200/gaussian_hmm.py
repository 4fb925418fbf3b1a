"""GaussianHMM -- kerehmm/gaussian_hmm.py:10-124 on the B200 engine."""
from __future__ import annotations

import numpy as np

from . import engine
from .abstract_hmm import AbstractHMM
from .distribution import GaussianDistribution


class GaussianHMM(AbstractHMM):
    """HMM with one Gaussian emission per state (gaussian_hmm.py:15-27).

    `GaussianHMM(number_of_states, dimensions, random_emissions=False, state_labels=None,
    random_transitions=..., upper_bounds=..., lower_bounds=..., verbose=...)`.  For
    `dimensions == 1` each state's `variance` attribute is used as the standard deviation
    (the reference passes it to scipy as `scale`); for `dimensions > 1` it is a full covariance.
    `emissionDistributions` is a plain list, as in the reference."""

    def __init__(self, number_of_states, dimensions, random_emissions=False, state_labels=None, *args, **kwargs):
        super(GaussianHMM, self).__init__(number_of_states, state_labels, *args, **kwargs)
        self.dimensions = dimensions
        if random_emissions:
            upper = kwargs.get('upper_bounds', [100 for _ in range(dimensions)])
            lower = kwargs.get('lower_bounds', [0 for _ in range(dimensions)])
            self.emissionDistributions = [GaussianDistribution(dimensions, random=True, upper_bounds=upper,
                                                               lower_bounds=lower) for _ in range(self.nStates)]
        else:
            self.emissionDistributions = [GaussianDistribution(dimensions) for _ in range(self.nStates)]

    def _params(self):
        pi = np.asarray(self.initialProbabilities, dtype=np.float64)
        A = np.asarray(self.transitionMatrix, dtype=np.float64)
        dims = self.emissionDistributions[0].dimensions
        if dims == 1:
            mean = np.array([float(np.asarray(d.mean).reshape(-1)[0]) for d in self.emissionDistributions])
            sd = np.array([float(np.asarray(d.variance).reshape(-1)[0]) for d in self.emissionDistributions])
            return engine.ModelParams(kind="gauss", pi=pi, A=A, mean=mean, sd=sd)
        means = np.stack([np.asarray(d.mean, dtype=np.float64) for d in self.emissionDistributions])
        covs = np.stack([np.asarray(d.variance, dtype=np.float64) for d in self.emissionDistributions])
        return engine.ModelParams(kind="gauss_full", pi=pi, A=A, mean=means, cov=covs)

    def _feature_dims(self):
        d = self.emissionDistributions[0].dimensions
        return 0 if d == 1 else d

    def estimate_initial_probabilities(self, gamma):
        """gaussian_hmm.py:29-32."""
        pi_new = np.zeros_like(self.initialProbabilities)
        pi_new[:] = gamma[0, :]
        return pi_new

    def do_pass(self, observations, verbose=False):
        """The reference's GaussianHMM.do_pass is stale log-space code that fails at its first
        lines at HEAD (`backward()` without the scale argument -> TypeError, gaussian_hmm.py:55-56;
        SURVEY.md Q9).  There are no semantics to reproduce; a Gaussian M-step is listed as a
        next-step extension."""
        raise TypeError("backward() missing 1 required positional argument: 'scale_coefficients'")
