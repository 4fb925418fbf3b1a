"""Emission distribution objects mirroring kerehmm/distribution.py.  They are the parameter
holders the HMM classes pack into device tables (`DiscreteDistribution.probabilities`,
`GaussianDistribution.mean/.variance`); `d[obs]` evaluates one probability on the host the way
the reference does.  The batched evaluation lives in the CUDA kernels."""
from __future__ import annotations

import numpy as np
from numpy.random import choice

from .util import random_simplex


class Distribution(object):
    def get_probability(self, observation, *args, **kwargs):
        raise NotImplementedError()

    def __getitem__(self, item, *args, **kwargs):
        return self.get_probability(item, *args, **kwargs)

    def emit(self):
        """Draw one observation."""
        raise NotImplementedError()


class DiscreteDistribution(Distribution):
    """Categorical distribution over `n` symbols (kerehmm/distribution.py:23-68)."""

    def __init__(self, n, randomize=False):
        self.n = n
        self.probabilities = np.array([1. / n] * n)
        if randomize:
            self.probabilities = random_simplex(n)

    @classmethod
    def from_probabilities(cls, probabilities):
        """A distribution holding a copy of `probabilities` (skips the uniform-table construction)."""
        d = cls.__new__(cls)
        d.probabilities = np.array(probabilities, dtype=np.float64)
        d.n = len(d.probabilities)
        return d

    @classmethod
    def from_row(cls, row):
        """A distribution whose table IS `row` (a float64 row view of a matrix the caller owns; no copy) -- the N
        distributions installed after a batched Baum-Welch pass share one (N, M) matrix this way."""
        d = cls.__new__(cls)
        d.probabilities = row
        d.n = len(row)
        return d

    def get_probability(self, observation, *args, **kwargs):
        return self.probabilities[observation]

    def emit(self):
        return choice(range(self.n), p=self.probabilities)

    def b_coefficient(self, other_dist):
        """Bhattacharyya coefficient as the reference defines it (distribution.py:41-64):
        sqrt(sum(p*q)).

        >>> d1, d2 = DiscreteDistribution(3), DiscreteDistribution(3)
        >>> bool(d1.b_coefficient(d2) == np.sqrt(np.sum([1. / 3 ** 2] * 3)))
        True
        """
        assert isinstance(other_dist, DiscreteDistribution)
        assert other_dist.n == self.n
        return np.sqrt(np.sum(self.probabilities * other_dist.probabilities))

    def b_distance(self, other_dist):
        return -np.log(self.b_coefficient(other_dist))


class GaussianDistribution(Distribution):
    """Gaussian emission (kerehmm/distribution.py:73-171).

    Univariate (`dimensions == 1`): `mean` and `variance` are scalars and -- exactly as in the
    reference, which hands `variance` to scipy as the *scale* -- `variance` acts as the standard
    deviation.  Multivariate: `mean` is (D,), `variance` a full (D, D) covariance.

    >>> m = GaussianDistribution(2)
    >>> m.mean.tolist(), m.variance.tolist()
    ([0.0, 0.0], [[1.0, 0.0], [0.0, 1.0]])
    >>> bool(np.isclose(m[(0, 0)], 0.15915494309189535))
    True
    >>> bool(np.isclose(GaussianDistribution(1, mean=3., variance=4.)[5.0], 0.08801633169107488))
    True
    """

    def __init__(self, dimensions, random=False, lower_bounds=0, upper_bounds=100,
                 mean=None, variance=None):
        self.dimensions = dimensions
        if random:
            draw = np.random.uniform(lower_bounds, upper_bounds, size=dimensions)
            self.mean = draw[0] if dimensions == 1 else draw
        elif mean is not None:
            if dimensions != 1 and np.shape(mean) != (dimensions,):
                raise ValueError("Mean has invalid shape: \nmean\t=\t{}".format(mean))
            self.mean = mean
        else:
            self.mean = 0. if dimensions == 1 else np.zeros(shape=(dimensions,))
        if variance is not None:
            self.variance = variance
        elif dimensions == 1:
            self.variance = 1.
        else:
            self.variance = np.eye(dimensions)

    def get_probability(self, observation, *args, **kwargs):
        x = np.asarray(observation, dtype=np.float64)
        if self.dimensions != 1:
            mean = np.asarray(self.mean, dtype=np.float64)
            L = np.linalg.cholesky(np.asarray(self.variance, dtype=np.float64))
            y = np.linalg.solve(L, (x - mean).reshape(-1, self.dimensions).T)
            expo = -0.5 * ((y * y).sum(axis=0) + 2.0 * np.log(np.diag(L)).sum()
                           + self.dimensions * np.log(2.0 * np.pi))
            out = np.exp(expo).reshape(x.shape[:-1])
            return float(out) if out.ndim == 0 else out
        z = (x - self.mean) / self.variance
        out = np.exp(-z * z / 2.0) / np.sqrt(2.0 * np.pi) / self.variance
        return float(out) if out.ndim == 0 else out

    def emit(self):
        if self.dimensions != 1:
            return np.random.multivariate_normal(self.mean, self.variance)
        return self.mean + self.variance * np.random.standard_normal()

    def __str__(self):
        return "\n            Gaussian(ndim={}, mean={}, covar={})\n            ".format(
            self.dimensions, self.mean, self.variance)

    __repr__ = __str__


class GaussianMixture(Distribution):
    """kerehmm/distribution.py:173-219 declares a mixture-of-Gaussians emission that no HMM class of the reference uses
    and whose constructor does not run on Python 3 (`distribution.py:202`); it is off the hot path
     and not provided -- the name exists so that code importing it fails with an explanation rather than
    an ImportError."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("GaussianMixture is unused by the reference's HMM classes and outside this engine's "
                                  "scope (SURVEY.md 2); GaussianDistribution covers Gaussian emissions")
