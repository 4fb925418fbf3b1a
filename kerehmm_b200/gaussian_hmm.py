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

    def estimate_transition_probabilities(self, *args, **kwargs):
        """gaussian_hmm.py:34-45 is part of the reference's stale log-space training code (it is only reachable from
        `do_pass`, which fails before calling it); the working path is `do_pass_batch` / `train_batch`."""
        raise NotImplementedError("GaussianHMM.estimate_transition_probabilities belongs to the reference's broken "
                                  "log-space do_pass; use do_pass_batch / train_batch")

    def estimate_emission_distributions(self, *args, **kwargs):
        """See `estimate_transition_probabilities` (gaussian_hmm.py:47-53)."""
        raise NotImplementedError("GaussianHMM.estimate_emission_distributions belongs to the reference's broken "
                                  "log-space do_pass; use do_pass_batch / train_batch")

    def do_pass(self, observations, verbose=False):
        """The reference's GaussianHMM.do_pass is stale log-space code that fails at its first
        lines at HEAD (`backward()` without the scale argument -> TypeError, gaussian_hmm.py:55-56;
        SURVEY.md Q9).  There are no semantics to reproduce; a Gaussian M-step is listed as a
        next-step extension."""
        raise TypeError("backward() missing 1 required positional argument: 'scale_coefficients'")

    # ------------------------------------------------- batched Baum-Welch (extension)
    def estep_batch(self, observations, lengths=None, dtype="float32", chunk=None, counts=None):
        """Textbook-EM sufficient statistics of a batch (1-D Gaussian emissions) as an
        `engine.GaussCounts` on the device.  EXTENSION: the reference's Gaussian training is broken at
        HEAD (see `do_pass`), so there are no reference semantics here; `DESIGN.md` states the definition."""
        ob = self._prep(observations, lengths)
        return engine.estep_gauss(self._params(), ob, self._real(dtype), counts=counts, chunk=chunk)

    def do_pass_batch(self, observations, lengths=None, dtype="float32", chunk=None, group=None, sd_min=1e-6):
        """One EM pass over a batch (this rank's shard under torch.distributed: the packed counts are
        all-reduced before the M-step).  Re-estimates pi, A (smoothed like the discrete path), means and
        standard deviations (`variance` attribute); returns the batch log-likelihood under the OLD
        parameters, which textbook EM never decreases."""
        from . import dist
        p = self._params()
        counts = self.estep_batch(observations, lengths, dtype, chunk)
        dist.all_reduce_counts(counts.buf, group=group)
        loglik = float(counts.buf[counts.sl["tail"]][0].item())
        pi, A, mean, sd = engine.mstep_gauss(p, counts, sd_min)
        self.initialProbabilities = pi
        self.transitionMatrix = A
        self.emissionDistributions = [GaussianDistribution(1, mean=float(m), variance=float(s)) for m, s in zip(mean, sd)]
        self.sanity_check()
        return loglik

    def train_batch(self, observations, lengths=None, iterations=10, dtype="float32", chunk=None, group=None,
                    tol=None, sd_min=1e-6):
        """`iterations` EM passes; returns the per-pass log-likelihoods (non-decreasing up to rounding)."""
        history = []
        for _ in range(iterations):
            history.append(self.do_pass_batch(observations, lengths, dtype, chunk, group, sd_min))
            if tol is not None and len(history) > 1 and history[-1] - history[-2] <= tol:
                break
        return history
