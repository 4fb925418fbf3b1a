#!/usr/bin/env python
"""Time the REAL keryil/kereHMM reference (converted by oracle/ref_loader.py, no arithmetic line touched) on this
container's CPU, on bounded samples of the BASELINE.json configurations, and write the numbers to
profiles/r1_reference_cpu_timing.json.  Test infrastructure: the reference cannot travel to the GPU box, so this is
run where /root/reference exists and the result is committed (bench.py --impl reference times the NumPy oracle PORT
on the box's cores instead; this file says how much slower the reference itself is).

    python oracle/time_reference.py
"""
import contextlib
import io
import json
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402
from oracle import hmm_oracle as O  # noqa: E402


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def timed(fn, min_s=0.5):
    n, t0 = 0, time.perf_counter()
    while True:
        with np.errstate(all="ignore"), quiet():
            fn()
        n += 1
        dt = time.perf_counter() - t0
        if dt >= min_s:
            return dt / n


def main():
    load_reference()
    from kerehmm.discrete_hmm import DiscreteHMM
    from kerehmm.gaussian_hmm import GaussianHMM
    out = {"host": platform.processor() or platform.machine(), "python": sys.version.split()[0], "numpy": np.__version__,
           "note": "single thread (the reference is pure Python over N-vectors); transitions = T * N^2 per pass",
           "cases": {}}

    def case(name, desc, N, T, passes, fn_ref, fn_port):
        s_ref, s_port = timed(fn_ref), timed(fn_port)
        tr = passes * T * N * N
        out["cases"][name] = {"workload": desc, "N": N, "T": T, "seconds_reference": s_ref, "seconds_oracle_port": s_port,
                              "transitions_per_s_reference": tr / s_ref, "transitions_per_s_oracle_port": tr / s_port}
        print(name, "reference %.4f s  port %.4f s  -> %.3e / %.3e transitions/s" % (s_ref, s_port, tr / s_ref, tr / s_port))

    # c1: the reference's own case, full size
    np.random.seed(1)
    h = DiscreteHMM(3, 4, random_transitions=True, random_emissions=True)
    obs = np.random.default_rng(1001).integers(0, 4, size=1000)
    pi, A = h.initialProbabilities.copy(), h.transitionMatrix.copy()
    Bm = np.array([d.probabilities for d in h.emissionDistributions])

    def c1_ref():
        a, s = h.forward(list(obs))
        h.backward(list(obs), scale_coefficients=s)
        h.viterbi_path(obs)

    def c1_port():
        E = O.emissions_discrete(Bm, obs)
        a, s = O.forward(pi, A, E)
        O.backward(A, E, s)
        O.viterbi(pi, A, Bm, obs[None])
    case("c1", "DiscreteHMM 3x4, forward + backward + Viterbi, one sequence T=1000 (full size)", 3, 1000, 3, c1_ref, c1_port)

    # c2: one sequence, T=128 of GaussianHMM N=8 forward + backward
    np.random.seed(2)
    g = GaussianHMM(8, 1, random_transitions=True, random_emissions=True)
    mean = np.array([float(np.ravel(d.mean)[0]) for d in g.emissionDistributions])
    sd = np.array([float(np.ravel(d.variance)[0]) for d in g.emissionDistributions])
    x = (mean[np.random.default_rng(1002).integers(0, 8, size=128)] + np.random.default_rng(7).standard_normal(128))
    gpi, gA = g.initialProbabilities.copy(), g.transitionMatrix.copy()

    def c2_ref():
        a, s = g.forward(list(x))
        g.backward(list(x), scale_coefficients=s)

    def c2_port():
        E = O.emissions_gaussian_1d(mean, sd, x)
        a, s = O.forward(gpi, gA, E)
        O.backward(gA, E, s)
    case("c2", "GaussianHMM N=8, forward + backward, one sequence T=128 (sample of 16384 x 2048)", 8, 128, 2, c2_ref, c2_port)

    # c3: one sequence, T=256 of DiscreteHMM 64x256 Viterbi
    np.random.seed(3)
    h3 = DiscreteHMM(64, 256, random_transitions=True, random_emissions=True)
    o3 = np.random.default_rng(1003).integers(0, 256, size=256)
    p3, A3 = h3.initialProbabilities.copy(), h3.transitionMatrix.copy()
    B3 = np.array([d.probabilities for d in h3.emissionDistributions])
    case("c3", "DiscreteHMM 64x256 Viterbi, one sequence T=256 (sample of 65536 x 4096)", 64, 256, 1,
         lambda: h3.viterbi_path(o3), lambda: O.viterbi(p3, A3, B3, o3[None]))

    # c4: one do_pass on one sequence, T=64 of DiscreteHMM 128x1024
    np.random.seed(4)
    h4 = DiscreteHMM(128, 1024, random_transitions=True, random_emissions=True)
    o4 = np.random.default_rng(1004).integers(0, 1024, size=64)
    p4, A4 = h4.initialProbabilities.copy(), h4.transitionMatrix.copy()
    B4 = np.array([d.probabilities for d in h4.emissionDistributions])

    def c4_ref():
        np.random.seed(4)
        hh = DiscreteHMM(128, 1024, random_transitions=True, random_emissions=True)
        hh.do_pass(list(o4))
    case("c4", "DiscreteHMM 128x1024 one Baum-Welch pass, one sequence T=64 (sample of 1M x 512; includes the constructor)",
         128, 64, 3, c4_ref, lambda: O.mstep_discrete(O.batch_stats_discrete(p4, A4, B4, o4[None])))

    path = os.path.join(ROOT, "profiles", "r1_reference_cpu_timing.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
