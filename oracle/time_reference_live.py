#!/usr/bin/env python
"""Time the REAL keryil/kereHMM reference (converted py2 -> py3 by oracle/ref_loader.py, no arithmetic line touched) on ONE
short sequence of a BASELINE.json configuration and print one JSON line.  Test infrastructure: called by bench.py's
cpu_baseline leg in a child interpreter (the reference's `kerehmm` package never shares a process with the product).

    python oracle/time_reference_live.py <c1|c2|c3|c4> <N> <M> <T> <seed>
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    kind, N, M, T, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    try:
        from oracle import ref_loader
        ref_loader.load_reference()
        from kerehmm.discrete_hmm import DiscreteHMM      # kerehmm/discrete_hmm.py (the reference itself)
        from kerehmm.gaussian_hmm import GaussianHMM
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            if kind == "c2":
                h = GaussianHMM(N, 1, random_transitions=True, random_emissions=True)
                mean = np.array([float(np.ravel(d.mean)[0]) for d in h.emissionDistributions])
                x = list(mean[np.random.default_rng(1000 + seed).integers(0, N, size=T)] +
                         np.random.default_rng(7).standard_normal(T))
                t0 = time.perf_counter()
                a, sc = h.forward(x)                                  # abstract_hmm.py:250-285
                h.backward(x, scale_coefficients=sc)                  # abstract_hmm.py:327-351
                passes, what = 2, "forward + backward"
            else:
                h = DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
                o = np.random.default_rng(1000 + seed).integers(0, M, size=T)
                t0 = time.perf_counter()
                if kind == "c1":
                    a, sc = h.forward(list(o))
                    h.backward(list(o), scale_coefficients=sc)
                    h.viterbi_path(o)                                 # abstract_hmm.py:203-239
                    passes, what = 3, "forward + backward + viterbi_path"
                elif kind == "c3":
                    h.viterbi_path(o)
                    passes, what = 1, "viterbi_path"
                else:
                    h.do_pass(list(o))                                # discrete_hmm.py:95-191
                    passes, what = 3, "do_pass (forward, backward, re-estimation)"
            dt = time.perf_counter() - t0
        out = {"transitions_per_s": passes * T * N * N / dt, "seconds": dt, "cores": 1,
               "sample": "%s of the reference class on ONE sequence of T=%d at the workload's N%s" % (
                   what, T, " and M" if M else ""),
               "where": "this box, this run (kerehmm converted py2 -> py3 by oracle/ref_loader.py, no arithmetic line touched)"}
    except Exception as e:
        out = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
