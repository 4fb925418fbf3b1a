"""Pin the CPU oracle (oracle/hmm_oracle.py) to the real reference.

Golden vectors come from the reference itself (oracle/make_golden.py); the known-answer
cases restate kerehmm/test/discrete_hmm_test.py:48-98.  CPU only."""
import numpy as np
import pytest

from oracle import hmm_oracle as O
from conftest import gcase

DISCRETE = ["D1", "D2", "D3", "D4"]
TOL = dict(rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("name", DISCRETE)
def test_forward_backward_gamma(golden, name):
    g = gcase(golden, name)
    E = O.emissions_discrete(g["B"], g["obs"])
    alpha, scale = O.forward(g["pi"], g["A"], E)
    beta = O.backward(g["A"], E, scale)
    np.testing.assert_allclose(alpha, g["alpha"], **TOL)
    np.testing.assert_allclose(scale, g["scale"], **TOL)
    np.testing.assert_allclose(beta, g["beta"], **TOL)
    np.testing.assert_allclose(O.gamma(alpha, beta, scale), g["gamma"], **TOL)
    np.testing.assert_allclose(O.forward_probability(scale), g["forward_probability"], rtol=1e-13)
    # gamma_ref rows sum to c_t
    np.testing.assert_allclose(g["gamma"].sum(axis=1), g["scale"], rtol=1e-12)


@pytest.mark.parametrize("name", DISCRETE + ["TIES", "KAT_l2r", "KAT_l2r_emis", "KAT_l2r_trans"])
def test_viterbi_bit_exact(golden, name):
    g = gcase(golden, name)
    with np.errstate(divide="ignore"):
        path, logp = O.viterbi(g["pi"], g["A"], g["B"], g["obs"])
    assert np.array_equal(path, g["path"])
    assert logp == g["logp"]          # bit-equal float64


def test_viterbi_batch_equals_single(golden):
    g = gcase(golden, "D2")
    rng = np.random.default_rng(5)
    obs = rng.integers(0, g["B"].shape[1], size=(9, 33))
    pb, lb = O.viterbi(g["pi"], g["A"], g["B"], obs)
    for b in range(obs.shape[0]):
        p1, l1 = O.viterbi(g["pi"], g["A"], g["B"], obs[b])
        assert np.array_equal(pb[b], p1) and lb[b] == l1


@pytest.mark.parametrize("name", ["D1", "D2", "D3"])
def test_zi_and_do_pass(golden, name):
    g = gcase(golden, name)
    E = O.emissions_discrete(g["B"], g["obs"])
    alpha, scale = O.forward(g["pi"], g["A"], E)
    beta = O.backward(g["A"], E, scale)
    np.testing.assert_allclose(O.xi_sum(g["A"], E, alpha, beta), g["zi_sum"], rtol=1e-11)
    if "zi" in g:
        z = O.zi(g["A"], E, alpha, beta)
        np.testing.assert_allclose(z, g["zi"], **TOL)
        # zi_ref[t].sum() == c_{t+1}
        np.testing.assert_allclose(z.sum(axis=(1, 2)), scale[1:], rtol=1e-12)
    pi2, A2, B2 = O.do_pass_discrete(g["pi"].copy(), g["A"].copy(), g["B"].copy(), g["obs"])
    np.testing.assert_allclose(pi2, g["pass_pi"], rtol=1e-12, atol=1e-22)
    np.testing.assert_allclose(A2, g["pass_A"], rtol=1e-11, atol=1e-22)
    np.testing.assert_allclose(B2, g["pass_B"], rtol=1e-11, atol=1e-22)


@pytest.mark.parametrize("name", ["D1", "D2"])
def test_train(golden, name):
    g = gcase(golden, name)
    pi, A, Bm, done = O.train_discrete(g["pi"].copy(), g["A"].copy(), g["B"].copy(), g["obs"],
                                       iterations=3, auto_stop=False)
    assert done == 4                                    # iterations+1 passes
    np.testing.assert_allclose(pi, g["train3_pi"], rtol=1e-9, atol=1e-20)
    np.testing.assert_allclose(A, g["train3_A"], rtol=1e-9, atol=1e-20)
    np.testing.assert_allclose(Bm, g["train3_B"], rtol=1e-9, atol=1e-20)
    pi, A, Bm, done = O.train_discrete(g["pi"].copy(), g["A"].copy(), g["B"].copy(), g["obs"],
                                       iterations=50, auto_stop=True)
    np.testing.assert_allclose(A, g["trainauto_A"], rtol=1e-8, atol=1e-18)
    np.testing.assert_allclose(Bm, g["trainauto_B"], rtol=1e-8, atol=1e-18)


def test_known_answer_uniform_forward(golden):
    """kerehmm/test/discrete_hmm_test.py:48-70 restated (np.product -> np.prod)."""
    n, m = 3, 3
    pt, pe = 1. / n, 1. / m
    pi, A, Bm = np.full(n, pt), np.full((n, n), pt), np.full((n, m), pe)
    fw, coefs = O.forward(pi, A, O.emissions_discrete(Bm, np.array([0, 0, 0])))
    fw1, _ = O.forward(pi, A, O.emissions_discrete(Bm, np.array([1, 1, 1])))
    assert np.array_equal(fw, fw1)
    line = np.array([pt * pe] * n)
    assert np.allclose(fw[0] * coefs[0], line, rtol=1e-15)
    line[:] = (line * pt).sum() * pe
    assert np.allclose(fw[1] * np.prod(coefs[:2]), line, rtol=1e-15)
    line[:] = (line * pt).sum() * pe
    assert np.allclose(fw[2] * np.prod(coefs[:3]), line, rtol=1e-15)
    g = gcase(golden, "KAT_uniform")
    np.testing.assert_allclose(fw, g["alpha"], rtol=1e-15)
    np.testing.assert_allclose(coefs, g["scale"], rtol=1e-15)
    # uniform model -> all-zero Viterbi path (first-index ties)
    path, logp = O.viterbi(pi, A, Bm, g["vit_obs"])
    assert np.array_equal(path, g["vit_path"]) and logp == g["vit_logp"]
    assert not path.any()


def test_known_answer_left_to_right(golden):
    """kerehmm/test/discrete_hmm_test.py:72-98: strict left-to-right -> path [0,1,2]."""
    pi, A = O.strict_left_to_right(3)
    g = gcase(golden, "KAT_l2r_emis")
    assert np.array_equal(pi, g["pi"]) and np.array_equal(A, g["A"])
    with np.errstate(divide="ignore"):
        path, _ = O.viterbi(pi, A, np.eye(3), np.array([0, 1, 2]))
    assert np.array_equal(path, [0, 1, 2])
    for obs in ([0, 0, 0], [1, 0, 1], [1, 1, 0]):
        path, _ = O.viterbi(pi, A, np.full((3, 3), 1 / 3.), np.array(obs))
        assert np.array_equal(path, [0, 1, 2])
    pi4, A4 = O.left_to_right(4)
    g4 = gcase(golden, "KAT_l2r")
    assert np.array_equal(pi4, g4["pi"]) and np.array_equal(A4, g4["A"])
    E = O.emissions_discrete(g4["B"], g4["obs"])
    a, c = O.forward(pi4, A4, E)
    np.testing.assert_allclose(a, g4["alpha"], **TOL)
    np.testing.assert_allclose(c, g4["scale"], **TOL)


@pytest.mark.parametrize("name", ["G1", "G2", "G3"])
def test_gaussian_forward_backward(golden, name):
    g = gcase(golden, name)
    E = O.emissions_gaussian_1d(g["mean"], g["sd"], g["obs"])
    if "emis" in g:
        np.testing.assert_allclose(E, g["emis"], rtol=1e-13, atol=1e-300)
    alpha, scale = O.forward(g["pi"], g["A"], E)
    beta = O.backward(g["A"], E, scale)
    np.testing.assert_allclose(alpha, g["alpha"], rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(scale, g["scale"], rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(beta, g["beta"], rtol=1e-10, atol=1e-300)
    if "gamma" in g:
        np.testing.assert_allclose(O.gamma(alpha, beta, scale), g["gamma"], rtol=1e-10, atol=1e-300)


def test_gaussian_pdf_against_scipy_and_reference(golden):
    from scipy.stats import multivariate_normal, norm
    g = gcase(golden, "G3")
    # "variance" is used as a standard deviation: mean 3, variance 4 at x=5
    assert np.isclose(O.gaussian_pdf_1d(5.0, 3.0, 4.0), g["pdf_3_4_at_5"], rtol=1e-15)
    assert np.isclose(g["pdf_3_4_at_5"], norm(loc=3., scale=4.).pdf(5.0), rtol=1e-15)
    x = np.linspace(-30, 30, 101)
    np.testing.assert_allclose(O.gaussian_pdf_1d(x, 1.5, 2.5), norm(loc=1.5, scale=2.5).pdf(x), rtol=1e-13, atol=1e-300)
    # doctest constants distribution.py:191-194 (bivariate standard normal)
    I = np.eye(2)
    assert np.isclose(O.gaussian_pdf_full(np.array([0., 0.]), np.zeros(2), I), 0.15915494309189535, rtol=1e-14)
    assert np.isclose(O.gaussian_pdf_full(np.array([1., 2.]), np.zeros(2), I), 0.013064233284684921, rtol=1e-14)
    gm = gcase(golden, "GM")
    for j in range(3):
        ref = multivariate_normal(mean=gm["means"][j], cov=gm["covs"][j]).pdf(gm["obs"])
        np.testing.assert_allclose(O.gaussian_pdf_full(gm["obs"], gm["means"][j], gm["covs"][j]), ref, rtol=1e-12)
    E = O.emissions_gaussian_full(gm["means"], gm["covs"], gm["obs"])
    alpha, scale = O.forward(gm["pi"], gm["A"], E)
    beta = O.backward(gm["A"], E, scale)
    np.testing.assert_allclose(alpha, gm["alpha"], rtol=1e-10)
    np.testing.assert_allclose(scale, gm["scale"], rtol=1e-10)
    np.testing.assert_allclose(beta, gm["beta"], rtol=1e-10)


def test_smooth_probabilities_and_simplex(golden):
    g = gcase(golden, "SMOOTH")
    x = g["in1"].copy()
    out = O.smooth_probabilities(x)
    assert out is x                                  # in place (util.py:57-60)
    assert np.array_equal(out, g["out1"])
    assert np.array_equal(O.smooth_probabilities(g["in2"].copy()), g["out2"])
    assert g["two_small_raises"] == 1
    with pytest.raises(ValueError):                  # 2 <= k < n small entries in a 2-D row
        O.smooth_probabilities(g["in3"].copy())
    with pytest.raises(ZeroDivisionError):
        O.smooth_probabilities(np.zeros(3))
    with pytest.raises(ValueError):
        O.smooth_probabilities(np.zeros((2, 2, 2)))
    s = gcase(golden, "SIMPLEX")
    np.random.seed(31)
    assert np.array_equal(O.random_simplex(7), s["one_d"])
    assert np.array_equal(O.random_simplex(5, two_d=True), s["two_d"])
    c = gcase(golden, "CTOR_D")
    np.random.seed(32)
    pi, A, Bm = O.random_discrete_model(4, 6)
    assert np.array_equal(pi, c["pi"]) and np.array_equal(A, c["A"]) and np.array_equal(Bm, c["B"])
    c = gcase(golden, "CTOR_G")
    np.random.seed(33)
    pi, A, mean, sd = O.random_gaussian_model(5)
    assert np.array_equal(pi, c["pi"]) and np.array_equal(A, c["A"])
    assert np.array_equal(mean, c["mean"]) and np.array_equal(sd, c["sd"])


def test_batch_stats_reduce_to_do_pass(golden):
    """Multi-sequence definition: with one sequence it IS do_pass."""
    g = gcase(golden, "D2")
    new = O.do_pass_discrete_batch(g["pi"], g["A"], g["B"], g["obs"][None])
    np.testing.assert_allclose(new[0], g["pass_pi"], rtol=1e-12, atol=1e-22)
    np.testing.assert_allclose(new[1], g["pass_A"], rtol=1e-11, atol=1e-22)
    np.testing.assert_allclose(new[2], g["pass_B"], rtol=1e-11, atol=1e-22)


def test_oracle_against_live_reference_if_present():
    """When /root/reference exists (build container), also compare on a fresh random case."""
    from oracle.ref_loader import load_reference, reference_available
    if not reference_available():
        pytest.skip("reference tree not present (GPU box)")
    import contextlib, io
    load_reference()
    from kerehmm.discrete_hmm import DiscreteHMM
    np.random.seed(99)
    h = DiscreteHMM(4, 5, random_transitions=True, random_emissions=True)
    obs = np.random.default_rng(99).integers(0, 5, size=35)
    pi, A = h.initialProbabilities.copy(), h.transitionMatrix.copy()
    Bm = np.stack([d.probabilities.copy() for d in h.emissionDistributions])
    path, logp = h.viterbi_path(obs)
    p2, l2 = O.viterbi(pi, A, Bm, obs)
    assert np.array_equal(path, p2) and logp == l2
    with contextlib.redirect_stdout(io.StringIO()):
        h.do_pass(list(obs))
    new = O.do_pass_discrete(pi, A, Bm, obs)
    np.testing.assert_allclose(new[1], h.transitionMatrix, rtol=1e-11)
    np.testing.assert_allclose(new[2], np.stack([d.probabilities for d in h.emissionDistributions]), rtol=1e-11)


# ------------------------------------------------------------------------------ sampling
@pytest.mark.parametrize("name", ["small", "wide"])
def test_simulate_restatement_matches_reference_golden(name):
    """`simulate` of the real reference (tests/golden/simulate_golden.npz, oracle/make_golden_simulate.py):
    with the MT19937 doubles the reference's numpy.random.choice calls consumed, the oracle's sampling
    procedure returns the reference's states and emissions exactly."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "simulate_golden.npz"))
    s, e = O.simulate_from_uniforms(g[name + "_pi"], g[name + "_A"], g[name + "_uniforms"], Bm=g[name + "_B"])
    assert np.array_equal(s, g[name + "_states"]) and np.array_equal(e, g[name + "_emissions"])


def test_gaussian_viterbi_extension_is_the_best_path():
    """EXTENSION, parity unpinned (the reference's viterbi_path fails on float observations, Q10): the oracle's
    Gaussian Viterbi must return the maximum over ALL state paths of log pi + sum log A + sum log N(x; mean, sd),
    with SciPy's logpdf as the independent statement of the density."""
    import itertools
    from scipy.stats import norm
    rng = np.random.default_rng(77)
    N, T = 3, 6
    np.random.seed(78)
    pi, A, mean, sd = O.random_gaussian_model(N)
    sd = sd * (0.5 + rng.random(N))
    x = mean[rng.integers(0, N, size=T)] + rng.standard_normal(T)
    path, logp = O.viterbi_gaussian_1d(pi, A, mean, sd, x)
    logE = np.stack([norm(mean[j], sd[j]).logpdf(x) for j in range(N)], axis=1)
    np.testing.assert_allclose(O.log_gaussian_pdf_1d(x, mean, sd), logE, rtol=1e-12, atol=1e-12)
    best, best_path = -np.inf, None
    for cand in itertools.product(range(N), repeat=T):
        s = np.log(pi[cand[0]]) + logE[0, cand[0]]
        for t in range(1, T):
            s += np.log(A[cand[t - 1], cand[t]]) + logE[t, cand[t]]
        if s > best:
            best, best_path = s, cand
    assert tuple(path) == best_path
    np.testing.assert_allclose(logp, best, rtol=1e-12)


def test_config1_full_length_golden():
    """BASELINE config 1 at its full length (3 states x 4 symbols, ONE sequence of T = 1000; tests/golden/c1_golden.npz from
    oracle/make_golden_c1.py): the oracle against the real reference's forward / backward / gamma / Viterbi outputs."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_golden.npz"))
    pi, A, Bm, obs = g["pi"], g["A"], g["B"], g["obs"]
    assert obs.shape == (1000,) and A.shape == (3, 3) and Bm.shape == (3, 4)
    E = O.emissions_discrete(Bm, obs)
    alpha, scale = O.forward(pi, A, E)
    beta = O.backward(A, E, scale)
    np.testing.assert_allclose(alpha, g["alpha"], rtol=1e-10)
    np.testing.assert_allclose(scale, g["scale"], rtol=1e-12)
    np.testing.assert_allclose(beta, g["beta"], rtol=1e-9)
    np.testing.assert_allclose(O.gamma(alpha, beta, scale), g["gamma"], rtol=1e-9)
    assert np.isclose(O.forward_probability(scale), g["forward_probability"], rtol=1e-12)
    with np.errstate(divide="ignore"):
        path, logp = O.viterbi(pi, A, Bm, obs)
    assert np.array_equal(path, g["path"]) and logp == g["logp"]          # bit-exact, 1000 steps
