"""GPU tests of the Gaussian Baum-Welch EXTENSION.  The reference's
GaussianHMM.do_pass raises at HEAD (Q9): there are no reference outputs to compare with, so the CUDA path is
checked against the oracle's statement of the same textbook-EM definition (whose forward / backward /
smoothing building blocks ARE pinned to the reference) and against EM's own invariant (the likelihood
never decreases)."""
import numpy as np
import pytest

from oracle import hmm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    import kerehmm_b200
    from kerehmm_b200 import engine
    engine.require_cuda()
    return kerehmm_b200


def gaussian_model(K, pi, A, mean, sd):
    h = K.GaussianHMM(A.shape[0], 1)
    h.initialProbabilities, h.transitionMatrix = pi.copy(), A.copy()
    h.emissionDistributions = [K.GaussianDistribution(1, mean=float(m), variance=float(s)) for m, s in zip(mean, sd)]
    return h


def make_case(N, B, T, seed):
    np.random.seed(seed)
    pi, A, mean, sd = O.random_gaussian_model(N, 0.0, min(10.0 * N, 100.0))
    sd = sd * np.random.uniform(0.7, 1.6, size=N)
    rng = np.random.default_rng(seed + 1)
    x = mean[rng.integers(0, N, size=(B, T))] + 1.3 * rng.standard_normal((B, T))
    lengths = rng.integers(max(2, T // 2), T + 1, size=B)
    return (pi, A, mean, sd), x, lengths


@pytest.mark.parametrize("N,B,T", [(5, 20, 60), (16, 9, 40), (130, 4, 12)])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_gauss_estep_and_do_pass_match_oracle(K, N, B, T, dtype):
    (pi, A, mean, sd), x, lengths = make_case(N, B, T, seed=80 + N)
    stats = O.batch_stats_gauss(pi, A, mean, sd, x, lengths)
    h = gaussian_model(K, pi, A, mean, sd)
    xin = x.astype(np.float32 if dtype == "float32" else np.float64)
    if dtype == "float32":       # the oracle must see what the device sees: float32 observations and parameters
        mean, sd = mean.astype(np.float32).astype(np.float64), sd.astype(np.float32).astype(np.float64)
        pi, A = pi.astype(np.float32).astype(np.float64), A.astype(np.float32).astype(np.float64)
        stats = O.batch_stats_gauss(pi, A, mean, sd, xin.astype(np.float64), lengths)
        h = gaussian_model(K, pi, A, mean, sd)
    c = h.estep_batch(xin, lengths=lengths, dtype=dtype, chunk=3).host()
    rtol = 3e-5 if dtype == "float32" else 1e-9
    np.testing.assert_allclose(c["pi"], stats["pi"], rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(c["xi_raw"] * A, stats["xi"], rtol=rtol, atol=1e-10 * stats["xi"].max())
    for k in ("gam", "s0", "s1", "s2"):
        np.testing.assert_allclose(c[k], stats[k], rtol=rtol, atol=1e-9 * np.abs(stats[k]).max(), err_msg=k)
    assert np.isclose(c["loglik"], stats["loglik"], rtol=1e-5) and c["nseq"] == B and c["flags"] == 0
    rpi, rA, rmean, rsd = O.mstep_gauss(stats)
    ll = h.do_pass_batch(xin, lengths=lengths, dtype=dtype)
    assert np.isclose(ll, stats["loglik"], rtol=1e-5)
    np.testing.assert_allclose(h.initialProbabilities, rpi, rtol=rtol, atol=1e-14)
    np.testing.assert_allclose(h.transitionMatrix, rA, rtol=rtol, atol=1e-14)
    got_mean = np.array([d.mean for d in h.emissionDistributions])
    got_sd = np.array([d.variance for d in h.emissionDistributions])
    np.testing.assert_allclose(got_mean, rmean, rtol=rtol, atol=1e-6 if dtype == "float32" else 1e-12)
    # variance = S2/S0 - mean^2 cancels for a state that explains ~one observation: compare variances with
    # an absolute slack proportional to mean^2 (the summation order differs between kernel and oracle)
    m2 = float(np.max(rmean * rmean))
    np.testing.assert_allclose(got_sd ** 2, rsd ** 2, rtol=2e-3 if dtype == "float32" else 1e-7,
                               atol=(1e-5 if dtype == "float32" else 1e-12) * m2)
    h.sanity_check()


def test_gauss_em_is_monotone_and_recovers_a_teacher(K):
    """Data from a teacher HMM drawn on the device (simulate_batch); EM from a perturbed start: the batch
    log-likelihood never decreases and the means move to the teacher's."""
    pi = np.array([0.5, 0.3, 0.2])
    A = np.array([[0.8, 0.15, 0.05], [0.1, 0.8, 0.1], [0.05, 0.15, 0.8]])
    tmean, tsd = np.array([-4.0, 0.5, 6.0]), np.array([0.7, 1.0, 1.4])
    teacher = gaussian_model(K, pi, A, tmean, tsd)
    x = teacher.simulate_batch(512, 200, seed=5).emissions
    h = gaussian_model(K, np.full(3, 1 / 3.0), np.full((3, 3), 1 / 3.0) * 0.7 + 0.1 * np.eye(3) * 3 / 1.0 / 1.0 * 1.0 / 1.0,
                       np.array([-2.0, 1.5, 4.0]), np.array([2.0, 2.0, 2.0]))
    h.transitionMatrix /= h.transitionMatrix.sum(axis=1, keepdims=True)
    hist = h.train_batch(x, iterations=15, dtype="float64")
    assert all(b >= a - 1e-6 * abs(a) for a, b in zip(hist, hist[1:])), hist
    got = np.sort([d.mean for d in h.emissionDistributions])
    np.testing.assert_allclose(got, tmean, atol=0.1)
    got_sd = np.array([d.variance for d in h.emissionDistributions])[np.argsort([d.mean for d in h.emissionDistributions])]
    np.testing.assert_allclose(got_sd, tsd, rtol=0.1)
