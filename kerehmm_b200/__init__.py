"""kerehmm_b200 -- B200-native batched HMM engine with the keryil/kereHMM API.

`AbstractHMM`, `DiscreteHMM`, `GaussianHMM` and the distribution objects keep the reference's
names and semantics; the recursions run as hand-written sm_100a CUDA kernels in libkhmm.so
(C ABI in include/khmm.h, bound with ctypes).  Importing the package is cheap; the first call that
needs the device loads the library and fails loudly if it (or a CUDA device) is missing.
"""
from .abstract_hmm import AbstractHMM, SimulationResult
from .discrete_hmm import DiscreteHMM
from .distribution import DiscreteDistribution, Distribution, GaussianDistribution
from .gaussian_hmm import GaussianHMM
from .util import CONVERGENCE_DELTA_LOG_LIKELIHOOD, DELTA_P, random_simplex, smooth_probabilities

__all__ = ["AbstractHMM", "DiscreteHMM", "GaussianHMM", "Distribution", "DiscreteDistribution",
           "GaussianDistribution", "SimulationResult", "smooth_probabilities", "random_simplex", "DELTA_P",
           "CONVERGENCE_DELTA_LOG_LIKELIHOOD"]
__version__ = "0.1.0"
