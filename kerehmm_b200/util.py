"""Host-side helpers mirroring kerehmm/util.py (constants, probability smoothing, random
simplex initialisation).  These run once per model construction / M-step on N- and M-sized
arrays; they are host logic of the reference API, not the accelerated path."""
from __future__ import annotations

from functools import reduce

import numpy as np

CONVERGENCE_DELTA_LOG_LIKELIHOOD = 1e-05   # kerehmm/util.py:3
DELTA_P = 1.0e-10                          # kerehmm/util.py:4


def add_logs(lst):
    """log(sum(exp(x))) over a list (kerehmm/util.py:7-13; `reduce` is a py2 builtin there)."""
    return reduce(np.logaddexp, lst)


def normalize(a):
    """Row/column alternating normaliser of kerehmm/util.py:16-20 (no callers in the reference)."""
    for i in range(len(a)):
        a[i, i:] = a[i, i:] / np.sum(a[i, i:]) * (1 - np.sum(a[i, :i]))
        a[i + 1:, i] = a[i + 1:, i] / np.sum(a[i + 1:, i]) * (1 - np.sum(a[:i + 1, i]))
    return a


def _smooth_row_2d(row):
    n = len(row)
    row /= row.sum()
    small = row < DELTA_P
    if small.any():
        z = row[small]
        # the reference uses the array of small VALUES as its counter (util.py:51-53); this only
        # broadcasts when exactly one (or every) entry is small and raises ValueError otherwise.
        row -= z * DELTA_P / (n - z)
        row[row < DELTA_P] = DELTA_P
    row /= row.sum()


def smooth_probabilities(probabilities):
    """Floor entries at DELTA_P and renormalise, in place, with the reference's exact 1-D / 2-D
    semantics including its error cases (kerehmm/util.py:39-63)."""
    if probabilities.ndim == 2:
        for i in range(probabilities.shape[0]):
            _smooth_row_2d(probabilities[i])
    elif probabilities.ndim == 1:
        n = len(probabilities)
        k = int(np.count_nonzero(probabilities < DELTA_P))
        if n == k:
            raise ZeroDivisionError("float division by zero")
        probabilities -= k * DELTA_P / (n - k)
        probabilities[probabilities < DELTA_P] = DELTA_P
        probabilities /= probabilities.sum()
    else:
        raise ValueError("Expected a 1d or 2d array, gotten shape {}.".format(probabilities.shape))
    return probabilities


def random_simplex(size, two_d=False, log_scale=False):
    """Random point(s) on the simplex: U(DELTA_P, 0.9) draws, then smoothing
    (kerehmm/util.py:23-36; same global-RNG draw order as the reference)."""
    n = size ** 2 if two_d else size
    arr = np.random.uniform(low=DELTA_P, high=.9, size=n)
    if two_d:
        arr = arr.reshape((size, size))
    arr = smooth_probabilities(arr)
    return np.log(arr) if log_scale else arr
