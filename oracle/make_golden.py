"""Generate tests/golden/reference_golden.npz by running the REAL reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

The reference is imported through `oracle/ref_loader.py` (py2->py3 token rewrite of
print/except only, in a temp dir; arithmetic untouched).  Inputs are seeded; both inputs
and the reference's outputs are stored, so the fixtures are self-contained and travel to
the GPU box.  Test infrastructure only.
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

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_golden.npz")


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def discrete_params(h):
    return (h.initialProbabilities.copy(), h.transitionMatrix.copy(),
            np.stack([d.probabilities.copy() for d in h.emissionDistributions]))


def main():
    load_reference()
    from kerehmm.discrete_hmm import DiscreteHMM
    from kerehmm.gaussian_hmm import GaussianHMM
    from kerehmm.distribution import GaussianDistribution
    from kerehmm import util as rutil

    G = {}

    # ---- random discrete models: every hot-path function --------------------------
    for name, (N, M, T, seed) in {"D1": (3, 4, 60, 11), "D2": (5, 7, 40, 12), "D3": (16, 32, 24, 13),
                                  "D4": (64, 256, 12, 14)}.items():
        np.random.seed(seed)
        h = DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
        obs = np.random.default_rng(1000 + seed).integers(0, M, size=T)
        pi, A, Bm = discrete_params(h)
        G[name + "/pi"], G[name + "/A"], G[name + "/B"], G[name + "/obs"] = pi, A, Bm, obs
        with np.errstate(all="ignore"), quiet():
            alpha, scale = h.forward(list(obs))
            beta = h.backward(list(obs), scale_coefficients=scale)
            gam = h.gamma(alpha, beta, scale)
            path, logp = h.viterbi_path(obs)
            fp = h.forward_probability(list(obs))
        G[name + "/alpha"], G[name + "/scale"], G[name + "/beta"], G[name + "/gamma"] = alpha, scale, beta, gam
        G[name + "/path"], G[name + "/logp"], G[name + "/forward_probability"] = path, np.float64(logp), np.float64(fp)
        if N <= 16:
            with quiet():
                z = h.zi(list(obs), alpha=alpha, beta=beta)
            G[name + "/zi_sum"] = z.sum(axis=0)
            if N <= 5:
                G[name + "/zi"] = z
            with quiet():
                h.do_pass(list(obs))
            p2 = discrete_params(h)
            G[name + "/pass_pi"], G[name + "/pass_A"], G[name + "/pass_B"] = p2
        if N <= 5:
            # train() from the ORIGINAL parameters (fresh model with the same seed)
            np.random.seed(seed)
            h2 = DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
            with quiet():
                h2.train(list(obs), iterations=3, auto_stop=False)
            p3 = discrete_params(h2)
            G[name + "/train3_pi"], G[name + "/train3_A"], G[name + "/train3_B"] = p3
            np.random.seed(seed)
            h3 = DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
            with quiet():
                h3.train(list(obs), iterations=50, auto_stop=True)
            p4 = discrete_params(h3)
            G[name + "/trainauto_pi"], G[name + "/trainauto_A"], G[name + "/trainauto_B"] = p4

    # ---- the reference's own known-answer tests (test/discrete_hmm_test.py:48-98) --
    h = DiscreteHMM(3, 3)
    fw, coefs = h.forward([0, 0, 0])
    G["KAT_uniform/alpha"], G["KAT_uniform/scale"] = fw, coefs
    path, logp = h.viterbi_path(np.array([2, 0, 1, 1]))
    G["KAT_uniform/vit_obs"], G["KAT_uniform/vit_path"], G["KAT_uniform/vit_logp"] = np.array([2, 0, 1, 1]), path, np.float64(logp)

    h = DiscreteHMM(3, 3)
    h.setup_strict_left_to_right(set_emissions=True)
    pi, A, Bm = discrete_params(h)
    obs = np.array([0, 1, 2])
    with np.errstate(all="ignore"):
        path, logp = h.viterbi_path(obs)
    G["KAT_l2r_emis/pi"], G["KAT_l2r_emis/A"], G["KAT_l2r_emis/B"], G["KAT_l2r_emis/obs"] = pi, A, Bm, obs
    G["KAT_l2r_emis/path"], G["KAT_l2r_emis/logp"] = path, np.float64(logp)

    h = DiscreteHMM(3, 3)
    h.setup_strict_left_to_right()
    pi, A, Bm = discrete_params(h)
    obs = np.array([1, 0, 1])
    path, logp = h.viterbi_path(obs)
    G["KAT_l2r_trans/pi"], G["KAT_l2r_trans/A"], G["KAT_l2r_trans/B"], G["KAT_l2r_trans/obs"] = pi, A, Bm, obs
    G["KAT_l2r_trans/path"], G["KAT_l2r_trans/logp"] = path, np.float64(logp)

    h = DiscreteHMM(4, 4)
    h.setup_left_to_right(set_emissions=True)
    pi, A, Bm = discrete_params(h)
    obs = np.array([0, 0, 1, 2, 2, 3, 0, 1])
    path, logp = h.viterbi_path(obs)
    alpha, scale = h.forward(list(obs))
    G["KAT_l2r/pi"], G["KAT_l2r/A"], G["KAT_l2r/B"], G["KAT_l2r/obs"] = pi, A, Bm, obs
    G["KAT_l2r/path"], G["KAT_l2r/logp"], G["KAT_l2r/alpha"], G["KAT_l2r/scale"] = path, np.float64(logp), alpha, scale

    # ---- exact ties in Viterbi (first-index rule) ----------------------
    N, M, T = 6, 4, 30
    h = DiscreteHMM(N, M)
    A = np.array([[.25, .25, .125, .125, .125, .125]] * N)
    A[3] = [.125, .125, .25, .25, .125, .125]
    h.transitionMatrix = A
    for i, d in enumerate(h.emissionDistributions):
        d.probabilities = np.array([.5, .25, .125, .125]) if i % 2 == 0 else np.array([.25, .5, .125, .125])
    obs = np.random.default_rng(77).integers(0, M, size=T)
    pi, A, Bm = discrete_params(h)
    path, logp = h.viterbi_path(obs)
    G["TIES/pi"], G["TIES/A"], G["TIES/B"], G["TIES/obs"] = pi, A, Bm, obs
    G["TIES/path"], G["TIES/logp"] = path, np.float64(logp)

    # ---- Gaussian models (forward/backward/gamma only: do_pass is broken, Q9) ------
    for name, (N, T, seed) in {"G1": (4, 30, 21), "G2": (8, 24, 22)}.items():
        np.random.seed(seed)
        g = GaussianHMM(N, 1, random_transitions=True, random_emissions=True)
        mean = np.array([d.mean for d in g.emissionDistributions], dtype=np.float64)
        sd = np.array([d.variance for d in g.emissionDistributions], dtype=np.float64)
        rng = np.random.default_rng(1000 + seed)
        obs = mean[rng.integers(0, N, size=T)] + rng.standard_normal(T)
        alpha, scale = g.forward(list(obs))
        beta = g.backward(list(obs), scale_coefficients=scale)
        gam = g.gamma(alpha, beta, scale)
        G[name + "/pi"], G[name + "/A"] = g.initialProbabilities.copy(), g.transitionMatrix.copy()
        G[name + "/mean"], G[name + "/sd"], G[name + "/obs"] = mean, sd, obs
        G[name + "/alpha"], G[name + "/scale"], G[name + "/beta"], G[name + "/gamma"] = alpha, scale, beta, gam
        G[name + "/forward_probability"] = np.float64(g.forward_probability(list(obs)))
        G[name + "/emis"] = np.array([g.emission_probability(None, o) for o in obs])

    # non-unit "variance" (used as sd) via explicit distributions
    np.random.seed(23)
    g = GaussianHMM(3, 1, random_transitions=True)
    g.emissionDistributions = [GaussianDistribution(1, mean=m, variance=v) for m, v in ((0., .5), (3., 4.), (-2., 1.5))]
    obs = np.random.default_rng(1023).normal(0, 3, size=20)
    alpha, scale = g.forward(list(obs))
    beta = g.backward(list(obs), scale_coefficients=scale)
    G["G3/pi"], G["G3/A"] = g.initialProbabilities.copy(), g.transitionMatrix.copy()
    G["G3/mean"], G["G3/sd"], G["G3/obs"] = np.array([0., 3., -2.]), np.array([.5, 4., 1.5]), obs
    G["G3/alpha"], G["G3/scale"], G["G3/beta"] = alpha, scale, beta
    G["G3/pdf_3_4_at_5"] = np.float64(GaussianDistribution(1, mean=3., variance=4.)[5.0])

    # multivariate (D=2, full covariance)
    np.random.seed(24)
    g = GaussianHMM(3, 2, random_transitions=True)
    means = np.array([[0., 0.], [2., -1.], [-3., 1.]])
    covs = np.array([[[1., .3], [.3, 2.]], [[.5, 0.], [0., .5]], [[2., -.4], [-.4, 1.]]])
    g.emissionDistributions = [GaussianDistribution(2, mean=means[i], variance=covs[i]) for i in range(3)]
    obs = np.random.default_rng(1024).normal(0, 2, size=(12, 2))
    alpha, scale = g.forward(list(obs))
    beta = g.backward(list(obs), scale_coefficients=scale)
    G["GM/pi"], G["GM/A"], G["GM/means"], G["GM/covs"], G["GM/obs"] = (
        g.initialProbabilities.copy(), g.transitionMatrix.copy(), means, covs, obs)
    G["GM/alpha"], G["GM/scale"], G["GM/beta"] = alpha, scale, beta

    # ---- util.smooth_probabilities / random_simplex / constructors -----------------
    x = np.array([0.5, 0.2, 1e-12, 0.3, 0.0])
    G["SMOOTH/in1"], G["SMOOTH/out1"] = x.copy(), rutil.smooth_probabilities(x.copy())
    x = np.array([[0.5, 0.2, 1e-12, 0.3], [1., 2., 3., 4.], [0.25, 0.25, 0.25, 0.25]])
    G["SMOOTH/in2"], G["SMOOTH/out2"] = x.copy(), rutil.smooth_probabilities(x.copy())
    x = np.array([[0.5, 1e-13, 1e-12, 0.5]])
    try:
        rutil.smooth_probabilities(x.copy())
        G["SMOOTH/two_small_raises"] = np.int64(0)
    except ValueError:
        G["SMOOTH/two_small_raises"] = np.int64(1)
    G["SMOOTH/in3"] = x
    np.random.seed(31)
    G["SIMPLEX/one_d"] = rutil.random_simplex(7)
    G["SIMPLEX/two_d"] = rutil.random_simplex(5, two_d=True)
    np.random.seed(32)
    h = DiscreteHMM(4, 6, random_transitions=True, random_emissions=True)
    G["CTOR_D/pi"], G["CTOR_D/A"], G["CTOR_D/B"] = discrete_params(h)
    np.random.seed(33)
    g = GaussianHMM(5, 1, random_transitions=True, random_emissions=True)
    G["CTOR_G/pi"], G["CTOR_G/A"] = g.initialProbabilities.copy(), g.transitionMatrix.copy()
    G["CTOR_G/mean"] = np.array([d.mean for d in g.emissionDistributions])
    G["CTOR_G/sd"] = np.array([d.variance for d in g.emissionDistributions], dtype=np.float64)

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, "with", len(G), "arrays,", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
