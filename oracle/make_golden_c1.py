"""Generate tests/golden/c1_golden.npz: BASELINE.json configs[0] (DiscreteHMM 3 states x 4 symbols, forward / backward /
Viterbi on ONE sequence of T = 1000) run through the REAL reference at full length.

    python oracle/make_golden_c1.py        (build container: needs /root/reference)

Same model and observation generators as bench.py's c1 workload (np.random.seed(1) for the reference's own constructor,
default_rng(1001) for the symbols).  Inputs and the reference's outputs are stored.  Test infrastructure only."""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "c1_golden.npz")


def main():
    load_reference()
    from kerehmm.discrete_hmm import DiscreteHMM
    np.random.seed(1)
    h = DiscreteHMM(3, 4, random_transitions=True, random_emissions=True)
    obs = np.random.default_rng(1001).integers(0, 4, size=1000)
    G = {"pi": h.initialProbabilities.copy(), "A": h.transitionMatrix.copy(),
         "B": np.stack([d.probabilities.copy() for d in h.emissionDistributions]), "obs": obs}
    with np.errstate(all="ignore"), contextlib.redirect_stdout(io.StringIO()):
        alpha, scale = h.forward(list(obs))
        beta = h.backward(list(obs), scale_coefficients=scale)
        gam = h.gamma(alpha, beta, scale)
        path, logp = h.viterbi_path(obs)
        fp = h.forward_probability(list(obs))
    G.update(alpha=alpha, scale=scale, beta=beta, gamma=gam, path=path, logp=np.float64(logp),
             forward_probability=np.float64(fp))
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, {k: getattr(v, "shape", ()) for k, v in G.items()})


if __name__ == "__main__":
    main()
