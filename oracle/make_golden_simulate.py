"""Generate tests/golden/simulate_golden.npz by running the REAL reference's `simulate`.

    python oracle/make_golden_simulate.py        (build container only: needs /root/reference)

`AbstractHMM.simulate` (abstract_hmm.py:458-469) draws through `numpy.random.choice`, which takes ONE
`random_sample()` double per call from the global MT19937 stream: for a discrete model the draws of step t
are (transition, emission) = doubles 2t and 2t+1 after `np.random.seed(s)`.  The fixture stores the model,
the seeds, those doubles and the reference's states/emissions, so that the sampling logic of the oracle
and of the CUDA kernel can be pinned to the reference exactly even though the device uses another
generator.  Test infrastructure only.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "simulate_golden.npz")


def main():
    load_reference()
    from kerehmm.discrete_hmm import DiscreteHMM
    out = {}
    for name, (N, M, T, nseq) in {"small": (3, 4, 40, 6), "wide": (16, 32, 25, 4)}.items():
        np.random.seed(900 + N)
        with contextlib.redirect_stdout(io.StringIO()):
            h = DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
        pi, A = h.initialProbabilities.copy(), h.transitionMatrix.copy()
        Bm = np.stack([d.probabilities.copy() for d in h.emissionDistributions])
        states, ems, U = [], [], []
        for k in range(nseq):
            seed = 7000 + 13 * k + N
            np.random.seed(seed)
            r = h.simulate(T)                       # reset=True: starts from pi
            states.append(np.array(r.states)); ems.append(np.array(r.emissions))
            d = np.random.RandomState(seed).random_sample(2 * T).reshape(T, 2)
            U.append(np.concatenate([d, np.zeros((T, 1))], axis=1))
        out.update({name + "_pi": pi, name + "_A": A, name + "_B": Bm, name + "_states": np.stack(states),
                    name + "_emissions": np.stack(ems), name + "_uniforms": np.stack(U)})
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
