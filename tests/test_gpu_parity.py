"""GPU parity: the CUDA path (through the Python classes -> ctypes -> libkhmm C ABI) against the
CPU oracle and the reference's golden vectors.  Integer results (Viterbi paths) must be bit-exact,
as must the float64 Viterbi log-probability; float64 recursions agree to ~1e-12, float32 ones to
the 1e-5 relative bound BASELINE.json states."""
import contextlib
import io

import numpy as np
import pytest

from conftest import gcase
from oracle import hmm_oracle as O

pytestmark = pytest.mark.gpu

F32_RTOL = 1e-5          # north_star: fp32 log-likelihoods / parameters within 1e-5 relative


@pytest.fixture(scope="module")
def K():
    import kerehmm_b200
    from kerehmm_b200 import engine
    engine.require_cuda()
    return kerehmm_b200


def discrete_model(K, pi, A, Bm):
    h = K.DiscreteHMM(A.shape[0], Bm.shape[1])
    h.initialProbabilities = pi.copy()
    h.transitionMatrix = A.copy()
    for i, d in enumerate(h.emissionDistributions):
        d.probabilities = Bm[i].copy()
    return h


def gaussian_model(K, pi, A, mean, sd):
    h = K.GaussianHMM(A.shape[0], 1)
    h.initialProbabilities = pi.copy()
    h.transitionMatrix = A.copy()
    h.emissionDistributions = [K.GaussianDistribution(1, mean=float(m), variance=float(s)) for m, s in zip(mean, sd)]
    return h


def random_discrete(K, N, M, seed):
    np.random.seed(seed)
    pi, A, Bm = O.random_discrete_model(N, M)
    return discrete_model(K, pi, A, Bm), (pi, A, Bm)


# ------------------------------------------------------------------ golden, single sequence
@pytest.mark.parametrize("name", ["D1", "D2", "D3", "D4"])
def test_single_sequence_api_matches_reference_golden(K, golden, name):
    g = gcase(golden, name)
    h = discrete_model(K, g["pi"], g["A"], g["B"])
    obs = [int(o) for o in g["obs"]]
    alpha, scale = h.forward(obs)
    assert alpha.dtype == np.float64 and alpha.shape == g["alpha"].shape
    np.testing.assert_allclose(alpha, g["alpha"], rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(scale, g["scale"], rtol=1e-11)
    beta = h.backward(obs, scale_coefficients=scale)
    np.testing.assert_allclose(beta, g["beta"], rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(h.gamma(alpha, beta, scale), g["gamma"], rtol=1e-11, atol=1e-300)
    assert np.isclose(h.forward_probability(obs), g["forward_probability"], rtol=1e-12)
    assert np.isclose(h.log_likelihood(obs), np.log(g["scale"]).sum(), rtol=1e-12)
    path, logp = h.viterbi_path(g["obs"])
    assert path.dtype == g["obs"].dtype                       # np.empty_like(observations), :234
    assert np.array_equal(path, g["path"])
    assert logp == g["logp"]                                  # bit-exact float64
    if "zi" in g:
        z = h.zi(obs, alpha=alpha, beta=beta)
        np.testing.assert_allclose(z, g["zi"], rtol=1e-11, atol=1e-300)
    with pytest.raises(TypeError):                            # inverted guards, SURVEY Q7
        h.zi(obs)


@pytest.mark.parametrize("name", ["D1", "D2", "D3"])
def test_do_pass_and_train_match_reference_golden(K, golden, name):
    g = gcase(golden, name)
    h = discrete_model(K, g["pi"], g["A"], g["B"])
    obs = [int(o) for o in g["obs"]]
    h.do_pass(obs)
    np.testing.assert_allclose(h.initialProbabilities, g["pass_pi"], rtol=1e-9, atol=1e-20)
    np.testing.assert_allclose(h.transitionMatrix, g["pass_A"], rtol=1e-9, atol=1e-20)
    np.testing.assert_allclose(h.emission_matrix(), g["pass_B"], rtol=1e-9, atol=1e-20)
    if "train3_A" in g:
        h = discrete_model(K, g["pi"], g["A"], g["B"])
        with contextlib.redirect_stdout(io.StringIO()) as out:
            h.train(obs, iterations=3, auto_stop=False)
        assert out.getvalue().count("ITERATION #") == 4       # iterations + 1 passes, SURVEY Q6
        np.testing.assert_allclose(h.transitionMatrix, g["train3_A"], rtol=1e-8, atol=1e-18)
        np.testing.assert_allclose(h.emission_matrix(), g["train3_B"], rtol=1e-8, atol=1e-18)
        np.testing.assert_allclose(h.initialProbabilities, g["train3_pi"], rtol=1e-8, atol=1e-18)
        h = discrete_model(K, g["pi"], g["A"], g["B"])
        with contextlib.redirect_stdout(io.StringIO()):
            h.train(obs, iterations=50, auto_stop=True)
        np.testing.assert_allclose(h.transitionMatrix, g["trainauto_A"], rtol=1e-7, atol=1e-16)


@pytest.mark.parametrize("name", ["TIES", "KAT_l2r", "KAT_l2r_emis", "KAT_l2r_trans"])
def test_viterbi_known_answers(K, golden, name):
    g = gcase(golden, name)
    h = discrete_model(K, g["pi"], g["A"], g["B"])
    path, logp = h.viterbi_path(g["obs"])
    assert np.array_equal(path, g["path"]) and logp == g["logp"]


def test_reference_known_answer_tests(K, golden):
    """kerehmm/test/discrete_hmm_test.py:48-98 run against the CUDA engine."""
    n = 3
    hmm = K.DiscreteHMM(n, n)
    pt, pe = 1. / n, 1. / n
    forward, coefs = hmm.forward([0, 0, 0])
    assert np.array_equal(forward, hmm.forward([1, 1, 1])[0])
    line = np.array([pt * pe] * n)
    assert np.allclose(forward[0] * coefs[0], line, rtol=1e-15)
    line[:] = (line * pt).sum() * pe
    assert np.allclose(forward[1] * np.prod(coefs[:2]), line, rtol=1e-15)
    line[:] = (line * pt).sum() * pe
    assert np.allclose(forward[2] * np.prod(coefs[:3]), line, rtol=1e-15)
    g = gcase(golden, "KAT_uniform")
    np.testing.assert_allclose(forward, g["alpha"], rtol=1e-15)
    path, _ = hmm.viterbi_path(g["vit_obs"])
    assert not path.any()                                      # uniform model: first-index ties -> zeros
    hmm = K.DiscreteHMM(n, n)
    hmm.setup_strict_left_to_right(set_emissions=True)
    path, _ = hmm.viterbi_path(np.array(range(n)))
    assert np.array_equal(path, np.array(range(n)))
    hmm = K.DiscreteHMM(n, n)
    hmm.setup_strict_left_to_right()
    path, _ = hmm.viterbi_path(np.random.randint(0, n - 1, n))
    assert np.array_equal(path, np.array(range(n)))


@pytest.mark.parametrize("name", ["G1", "G2", "G3"])
def test_gaussian_single_sequence_golden(K, golden, name):
    g = gcase(golden, name)
    h = gaussian_model(K, g["pi"], g["A"], g["mean"], g["sd"])
    obs = list(g["obs"])
    alpha, scale = h.forward(obs)
    np.testing.assert_allclose(alpha, g["alpha"], rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(scale, g["scale"], rtol=1e-9, atol=1e-300)
    beta = h.backward(obs, scale_coefficients=scale)
    np.testing.assert_allclose(beta, g["beta"], rtol=1e-9, atol=1e-300)
    if "emis" in g:
        np.testing.assert_allclose(h.emission_probability(None, obs[3]), g["emis"][3], rtol=1e-13)
    with pytest.raises(TypeError):                            # GaussianHMM.do_pass is broken upstream (Q9)
        h.do_pass(obs)
    with pytest.raises(IndexError):                           # Viterbi on float observations (Q10)
        h.viterbi_path(np.array(obs))


def test_multivariate_gaussian_golden(K, golden):
    g = gcase(golden, "GM")
    h = K.GaussianHMM(3, 2)
    h.initialProbabilities, h.transitionMatrix = g["pi"].copy(), g["A"].copy()
    h.emissionDistributions = [K.GaussianDistribution(2, mean=g["means"][i], variance=g["covs"][i]) for i in range(3)]
    alpha, scale = h.forward(g["obs"])
    np.testing.assert_allclose(alpha, g["alpha"], rtol=1e-9)
    np.testing.assert_allclose(scale, g["scale"], rtol=1e-9)
    np.testing.assert_allclose(h.backward(g["obs"], scale), g["beta"], rtol=1e-9)


# ------------------------------------------------------------------------ batched Viterbi
@pytest.mark.parametrize("N,M,B,T,dtype", [(3, 4, 70, 100, np.int64), (64, 256, 40, 64, np.uint8),
                                           (128, 1024, 9, 40, np.uint16), (300, 16, 5, 30, np.int32),
                                           (17, 5, 33, 1, np.uint8), (32, 16, 100, 50, np.uint8),
                                           (64, 256, 70, 64, np.uint8), (128, 1024, 65, 40, np.uint16)])
def test_viterbi_batch_bit_exact(K, N, M, B, T, dtype):
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=N + M)
    obs = np.random.default_rng(N).integers(0, M, size=(B, T)).astype(dtype)
    path, logp = h.viterbi_batch(obs)
    with np.errstate(divide="ignore"):
        rp, rl = O.viterbi(pi, A, Bm, obs.astype(np.int64))
    assert np.array_equal(path, rp)
    assert np.array_equal(logp, rl)                            # bit-exact float64


def test_viterbi_batch_ragged_and_cuda_tensor_input(K):
    import torch
    N, M, B, T = 20, 11, 37, 50
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=5)
    rng = np.random.default_rng(6)
    obs = rng.integers(0, M, size=(B, T)).astype(np.uint8)
    lengths = rng.integers(1, T + 1, size=B)
    path, logp = h.viterbi_batch(torch.from_numpy(obs).cuda(), lengths=lengths, path_dtype=torch.int32)
    assert path.is_cuda and logp.is_cuda                       # CUDA in -> CUDA out
    path, logp = path.cpu().numpy(), logp.cpu().numpy()
    for b in range(B):
        L = lengths[b]
        rp, rl = O.viterbi(pi, A, Bm, obs[b, :L].astype(np.int64))
        assert np.array_equal(path[b, :L], rp) and logp[b] == rl
        assert not path[b, L:].any()


@pytest.mark.parametrize("N", [32, 64, 128])
def test_viterbi_register_tile_kernels_ragged(K, N):
    """B >= 64 with N in {32, 64, 128} takes the register-tile kernel; ragged lengths, exact ties
    (quantised probabilities) and zero probabilities included."""
    M, B, T = 12, 97, 37
    rng = np.random.default_rng(N)
    A = rng.integers(0, 4, size=(N, N)).astype(np.float64)      # many exact ties and zeros
    A[np.arange(N), np.arange(N)] += 1
    A /= A.sum(axis=1, keepdims=True)
    Bm = rng.integers(1, 4, size=(N, M)).astype(np.float64)
    Bm /= Bm.sum(axis=1, keepdims=True)
    pi = np.full(N, 1.0 / N)
    h = discrete_model(K, pi, A, Bm)
    obs = rng.integers(0, M, size=(B, T)).astype(np.uint8)
    lengths = rng.integers(1, T + 1, size=B)
    path, logp = h.viterbi_batch(obs, lengths=lengths)
    with np.errstate(divide="ignore"):
        for b in range(B):
            L = lengths[b]
            rp, rl = O.viterbi(pi, A, Bm, obs[b, :L].astype(np.int64))
            assert np.array_equal(path[b, :L], rp), b
            assert logp[b] == rl


@pytest.mark.parametrize("kind", ["near_ties", "uniform", "left_to_right", "long", "unnormalised", "banded"])
@pytest.mark.parametrize("N", [32, 64, 128])
def test_viterbi_screened_kernel_adversarial(K, N, kind):
    """The float32-screened kernel (N in {32, 64, 128}, B >= 64) must stay bit-exact where float32 cannot decide:
    transition rows that differ in the 9th digit only, exact ties everywhere, -inf rows and columns, |delta| ~ 1e4
    (long sequences), log A > 0 (unnormalised rows) and banded matrices (most candidates -inf)."""
    M, B, T = 16, 70, 60
    rng = np.random.default_rng(1000 + N)
    A = rng.random((N, N)) + 0.1
    A /= A.sum(axis=1, keepdims=True)
    Bm = rng.random((N, M)) + 0.05
    Bm /= Bm.sum(axis=1, keepdims=True)
    pi = rng.random(N)
    pi /= pi.sum()
    if kind == "near_ties":
        base = rng.random(N) + 0.5
        A = np.tile(base / base.sum(), (N, 1)) * (1.0 + 1e-9 * rng.standard_normal((N, N)))   # float32-identical columns
        Bm = np.tile(Bm[0], (N, 1)) * (1.0 + 1e-10 * rng.standard_normal((N, M)))
    elif kind == "uniform":
        A = np.full((N, N), 1.0 / N)
        Bm = np.full((N, M), 1.0 / M)
        pi = np.full(N, 1.0 / N)
    elif kind == "left_to_right":
        A = np.zeros((N, N))
        for i in range(N):
            A[i, i] = 0.5
            A[i, min(i + 1, N - 1)] += 0.5
        pi = np.zeros(N)
        pi[0] = 1.0
    elif kind == "long":
        T = 3000 if N <= 64 else 1200
        B = 64
    elif kind == "unnormalised":
        A = A * 50.0                                             # log A > 0 for many entries
    elif kind == "banded":
        mask = np.abs(np.subtract.outer(np.arange(N), np.arange(N))) <= 2
        A = np.where(mask, A, 0.0)
        A /= A.sum(axis=1, keepdims=True)
    h = discrete_model(K, pi, A, Bm)
    obs = rng.integers(0, M, size=(B, T)).astype(np.uint8)
    path, logp = h.viterbi_batch(obs)
    with np.errstate(divide="ignore"):
        rp, rl = O.viterbi(pi, A, Bm, obs.astype(np.int64))
    assert np.array_equal(path, rp)
    assert np.array_equal(logp, rl)


@pytest.mark.parametrize("N,B,T,dtype", [(3, 5, 40, np.float64), (8, 70, 100, np.float32), (64, 66, 30, np.float32),
                                          (300, 4, 25, np.float64), (32, 64, 50, np.float64), (128, 70, 40, np.float32),
                                          (64, 200, 300, np.float64)])
def test_viterbi_batch_gaussian_extension(K, N, B, T, dtype):
    """EXTENSION: batched Viterbi for 1-D Gaussian emissions, bit-exact against the oracle's
    statement of the same definition (ragged lengths included); the single-sequence reference method keeps failing
    on float observations like the reference (Q10)."""
    np.random.seed(300 + N)
    pi, A, mean, sd = O.random_gaussian_model(N)
    sd = sd * (0.5 + np.random.random(N))                       # the reference initialises every sd to 1
    h = gaussian_model(K, pi, A, mean, sd)
    rng = np.random.default_rng(301 + N)
    x = (mean[rng.integers(0, N, size=(B, T))] + 2.0 * rng.standard_normal((B, T))).astype(dtype)
    lengths = rng.integers(1, T + 1, size=B)
    path, logp = h.viterbi_batch(x, lengths=lengths)
    for b in range(B):
        L = lengths[b]
        rp, rl = O.viterbi_gaussian_1d(pi, A, mean, sd, x[b, :L])
        assert np.array_equal(path[b, :L], rp), b
        assert logp[b] == rl
        assert not path[b, L:].any()
    with pytest.raises(IndexError):
        h.viterbi_path(x[0])


def test_viterbi_with_zero_probabilities(K):
    """log(0) = -inf entries; -inf columns resolve to index 0."""
    N, M = 5, 4
    A = np.zeros((N, N))
    for i in range(N):
        A[i, (i + 1) % N] = 0.75
        A[i, i] = 0.25
    Bm = np.full((N, M), 0.25)
    Bm[2] = [1, 0, 0, 0]
    pi = np.array([1., 0, 0, 0, 0])
    h = discrete_model(K, pi, A, Bm)
    obs = np.random.default_rng(3).integers(0, M, size=(16, 25))
    path, logp = h.viterbi_batch(obs)
    with np.errstate(divide="ignore"):
        rp, rl = O.viterbi(pi, A, Bm, obs)
    assert np.array_equal(path, rp) and np.array_equal(logp, rl)


# ----------------------------------------------------------------- batched forward/backward
@pytest.mark.parametrize("N,M,B,T", [(3, 4, 65, 120), (8, 16, 33, 90), (40, 64, 10, 50), (128, 1024, 6, 40),
                                     (260, 32, 3, 20)])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_forward_backward_batch_discrete(K, N, M, B, T, dtype):
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=N)
    rng = np.random.default_rng(N + 1)
    obs = rng.integers(0, M, size=(B, T))
    lengths = rng.integers(max(1, T // 2), T + 1, size=B)
    alpha, scale, ll = h.forward_batch(obs, lengths=lengths, dtype=dtype)
    beta = h.backward_batch(obs, scale, lengths=lengths, dtype=dtype)
    gam, ll2 = h.posterior_batch(obs, lengths=lengths, dtype=dtype)
    rtol = 2e-5 if dtype == "float32" else 1e-10
    for b in range(B):
        L = lengths[b]
        E = O.emissions_discrete(Bm, obs[b, :L])
        ra, rc = O.forward(pi, A, E)
        rb = O.backward(A, E, rc)
        np.testing.assert_allclose(alpha[b, :L], ra, rtol=rtol, atol=1e-30)
        np.testing.assert_allclose(scale[b, :L], rc, rtol=rtol)
        np.testing.assert_allclose(beta[b, :L], rb, rtol=rtol * 5, atol=1e-30)
        np.testing.assert_allclose(gam[b, :L], ra * rb, rtol=rtol * 5, atol=1e-30)
        assert np.isclose(ll[b], np.log(rc).sum(), rtol=F32_RTOL if dtype == "float32" else 1e-12)
        assert ll2[b] == ll[b]


@pytest.mark.parametrize("N,B,T", [(8, 300, 257), (3, 64, 100), (12, 40, 64), (16, 33, 50), (1, 5, 10)])
def test_fused_loglik_small_gaussian(K, N, B, T):
    np.random.seed(100 + N)
    pi, A, mean, sd = O.random_gaussian_model(N)
    sd = sd * np.random.uniform(0.5, 2.0, size=N)
    h = gaussian_model(K, pi, A, mean, sd)
    rng = np.random.default_rng(N)
    obs = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
    lengths = rng.integers(1, T + 1, size=B)
    ll, llb = h.log_likelihood_batch(obs, lengths=lengths)
    _, _, ll_generic = h.forward_batch(obs, lengths=lengths, return_alpha=False)
    for b in range(B):
        E = O.emissions_gaussian_1d(mean, sd, obs[b, :lengths[b]].astype(np.float64))
        ref = np.log(O.forward(pi, A, E)[1]).sum()
        assert np.isclose(ll[b], ref, rtol=F32_RTOL, atol=1e-4), (b, ll[b], ref)
        assert np.isclose(llb[b], ref, rtol=F32_RTOL, atol=1e-4), (b, llb[b], ref)
        assert np.isclose(ll_generic[b], ref, rtol=F32_RTOL, atol=1e-4)


def test_fused_loglik_small_discrete(K):
    N, M, B, T = 3, 4, 50, 1000
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=1)
    obs = np.random.default_rng(2).integers(0, M, size=(B, T)).astype(np.uint8)
    ll, llb = h.log_likelihood_batch(obs)
    ref = O.loglik(O.forward(pi, A, O.emissions_discrete(Bm, obs))[1])
    np.testing.assert_allclose(ll, ref, rtol=F32_RTOL)
    np.testing.assert_allclose(llb, ref, rtol=F32_RTOL)


@pytest.mark.parametrize("N", [5, 6, 7, 8])
def test_fused_loglik_warp_tile_kernel(K, N):
    """N in 5..8 takes the kernel that runs the matrix-vector products of 16 sequences as one warp-level tensor-core tile
    (split tf32 operands): float64 oracle within the fp32 bound, Gaussian and discrete, ragged batches whose size is not
    a multiple of the tile, lengths from 1 up; and agreement with the thread-per-sequence kernel it replaces."""
    from kerehmm_b200 import _lib
    B, T, M = 117, 203, 11
    np.random.seed(300 + N)
    pi, A, mean, sd = O.random_gaussian_model(N)
    sd = sd * np.random.uniform(0.5, 2.0, size=N)
    hg = gaussian_model(K, pi, A, mean, sd)
    hd, (pid, Ad, Bm) = random_discrete(K, N, M, seed=310 + N)
    rng = np.random.default_rng(320 + N)
    x = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
    sym = rng.integers(0, M, size=(B, T)).astype(np.uint8)
    lengths = rng.integers(1, T + 1, size=B)
    lengths[:20] = T                   # whole tiles on the unrolled path
    lengths[40] = 1
    ref_g = np.array([np.log(O.forward(pi, A, O.emissions_gaussian_1d(mean, sd, x[b, :lengths[b]].astype(np.float64)))[1]).sum()
                      for b in range(B)])
    ref_d = np.array([np.log(O.forward(pid, Ad, O.emissions_discrete(Bm, sym[b, :lengths[b]]))[1]).sum() for b in range(B)])
    got = {}
    try:
        for sw in (1, 0):
            _lib.call("khmm_debug_switch", b"KHMM_SMALL_MMA", sw)
            got[sw] = (hg.log_likelihood_batch(x, lengths=lengths), hd.log_likelihood_batch(sym, lengths=lengths))
    finally:
        _lib.call("khmm_debug_switch", b"KHMM_SMALL_MMA", -1)
    for sw in (1, 0):
        (ll, llb), (dl, dlb) = got[sw]
        for r in (ll, llb):
            np.testing.assert_allclose(r, ref_g, rtol=F32_RTOL, atol=1e-4)
        for r in (dl, dlb):
            np.testing.assert_allclose(r, ref_d, rtol=F32_RTOL, atol=1e-4)
    np.testing.assert_allclose(got[1][0][0], got[0][0][0], rtol=2e-6, atol=2e-5)
    np.testing.assert_allclose(got[1][1][0], got[0][1][0], rtol=2e-6, atol=2e-5)
    # whole sequences of a multiple of 8 steps take the vector-load instantiation (switch value 2: scalar loads)
    xw = np.ascontiguousarray(x[:, :200])
    ref_w = np.array([np.log(O.forward(pi, A, O.emissions_gaussian_1d(mean, sd, xw[b].astype(np.float64)))[1]).sum()
                      for b in range(B)])
    try:
        vec = hg.log_likelihood_batch(xw)
        _lib.call("khmm_debug_switch", b"KHMM_SMALL_MMA", 2)
        scalar = hg.log_likelihood_batch(xw)
    finally:
        _lib.call("khmm_debug_switch", b"KHMM_SMALL_MMA", -1)
    for r in vec + scalar:
        np.testing.assert_allclose(r, ref_w, rtol=F32_RTOL, atol=1e-4)
    np.testing.assert_allclose(vec[0], scalar[0], rtol=2e-6, atol=2e-5)


@pytest.mark.parametrize("N,B,T", [(128, 200, 40), (256, 129, 25), (1024, 64, 12)])
def test_tensor_core_loglik_gaussian(K, N, B, T):
    """tcgen05 (tf32) recursion: forward and backward log-likelihoods against the float64 oracle.
    Stated bound for this reduced-precision path: 2e-5 relative (measured ~1e-6)."""
    np.random.seed(200 + N)
    pi, A, mean, sd = O.random_gaussian_model(N)
    sd = sd * np.random.uniform(0.7, 1.5, size=N)
    h = gaussian_model(K, pi, A, mean, sd)
    rng = np.random.default_rng(N)
    obs = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
    ll, llb = h.log_likelihood_batch(obs)
    assert llb is not None                                   # the tensor-core path returns both directions
    ll_simt, none = h.log_likelihood_batch(obs, tensor_cores=False)
    assert none is None
    E = O.emissions_gaussian_1d(mean, sd, obs.astype(np.float64))
    ref = O.loglik(O.forward(pi, A, E)[1])
    np.testing.assert_allclose(ll_simt, ref, rtol=F32_RTOL)
    np.testing.assert_allclose(ll, ref, rtol=2e-5)
    np.testing.assert_allclose(llb, ref, rtol=2e-5)


def test_tensor_core_loglik_discrete(K):
    N, M, B, T = 128, 64, 140, 30
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=77)
    obs = np.random.default_rng(78).integers(0, M, size=(B, T)).astype(np.uint8)
    ll, llb = h.log_likelihood_batch(obs)
    ref = O.loglik(O.forward(pi, A, O.emissions_discrete(Bm, obs))[1])
    np.testing.assert_allclose(ll, ref, rtol=2e-5)
    np.testing.assert_allclose(llb, ref, rtol=2e-5)
    # ragged batches stay on the tensor-core kernel (discrete table gather; the padding may hold any symbol)
    lengths = np.full(B, T); lengths[3] = T - 5; lengths[7] = 1
    obs2 = obs.copy(); obs2[3, T - 5:] = 255; obs2[7, 1:] = 200
    ll2, llb2 = h.log_likelihood_batch(obs2, lengths=lengths)
    ref2 = ref.copy()
    for b in (3, 7):
        ref2[b] = O.loglik(O.forward(pi, A, O.emissions_discrete(Bm, obs[b, :lengths[b]]))[1])
    np.testing.assert_allclose(ll2, ref2, rtol=2e-5, atol=3e-3)
    np.testing.assert_allclose(llb2, ref2, rtol=2e-5, atol=3e-3)
    ll3, none = h.log_likelihood_batch(obs2, lengths=lengths, tensor_cores=False)      # float32 SIMT forward kernel
    assert none is None
    np.testing.assert_allclose(ll3, ref2, rtol=F32_RTOL)


def test_gaussian_forward_backward_batch_generic(K):
    N, B, T = 8, 25, 80
    np.random.seed(3)
    pi, A, mean, sd = O.random_gaussian_model(N)
    h = gaussian_model(K, pi, A, mean, sd)
    rng = np.random.default_rng(4)
    obs = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T)))
    for dtype, rtol in (("float64", 1e-9), ("float32", 1e-4)):
        alpha, scale, ll = h.forward_batch(obs.astype(np.float32 if dtype == "float32" else np.float64), dtype=dtype)
        o = obs.astype(np.float32).astype(np.float64) if dtype == "float32" else obs
        E = O.emissions_gaussian_1d(mean, sd, o)
        ra, rc = O.forward(pi, A, E)
        np.testing.assert_allclose(scale, rc, rtol=rtol, atol=1e-300)
        np.testing.assert_allclose(alpha, ra, rtol=rtol, atol=1e-12)
        np.testing.assert_allclose(ll, np.log(rc).sum(axis=1), rtol=F32_RTOL)


# ----------------------------------------------------------------------------- Baum-Welch
@pytest.mark.parametrize("N,M,B,T", [(3, 4, 12, 60), (16, 32, 9, 40), (128, 64, 5, 24), (130, 9, 3, 12)])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_do_pass_batch_matches_oracle(K, N, M, B, T, dtype):
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=10 + N)
    rng = np.random.default_rng(11)
    obs = rng.integers(0, M, size=(B, T))
    lengths = rng.integers(max(2, T // 2), T + 1, size=B)
    stats = None
    for b in range(B):
        s = O.sequence_stats_discrete(pi, A, Bm, obs[b, :lengths[b]])
        stats = s if stats is None else {k: stats[k] + s[k] for k in s}
    rpi, rA, rB = O.mstep_discrete(stats)
    # E-step statistics first (chunked so that the chunk loop is exercised)
    c = h.estep_batch(obs, lengths=lengths, dtype=dtype, chunk=4).host()
    rtol = 2e-5 if dtype == "float32" else 1e-10
    np.testing.assert_allclose(c["pi"], stats["pi"], rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(c["xi_raw"] * A, stats["xi"], rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(c["gam"], stats["gam"], rtol=rtol)
    np.testing.assert_allclose(c["emis_num"], stats["emis_num"], rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(c["emis_den"], stats["emis_den"], rtol=rtol)
    assert np.isclose(c["loglik"], stats["loglik"], rtol=F32_RTOL) and c["nseq"] == B and c["flags"] == 0
    ll = h.do_pass_batch(obs, lengths=lengths, dtype=dtype)
    assert np.isclose(ll, stats["loglik"], rtol=F32_RTOL)
    np.testing.assert_allclose(h.initialProbabilities, rpi, rtol=rtol, atol=1e-14)
    np.testing.assert_allclose(h.transitionMatrix, rA, rtol=rtol, atol=1e-14)
    np.testing.assert_allclose(h.emission_matrix(), rB, rtol=rtol, atol=1e-14)


# tf32 tensor-core E-step: every (sequence, step) term carries an unbiased ~5e-4 relative rounding
# error (10-bit mantissa operands), so single entries of the statistics are compared at 1e-2 and
# sums over many terms much tighter; DESIGN.md 3.5 states the bound.
TC_TERM_RTOL = 1e-2


@pytest.mark.parametrize("M,B,T,odt", [(64, 300, 40, np.int64), (1024, 128, 33, np.uint16), (9, 77, 2, np.uint8),
                                       (16, 130, 1, np.int32)])
def test_tensor_core_estep_matches_oracle(K, M, B, T, odt):
    """Fused tcgen05 E-step (N = 128): forward stash, scale factors and all reference-weighted
    statistics against the float64 oracle, including a partial last tile (B % 128 != 0)."""
    from kerehmm_b200 import engine
    N = 128
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=70 + M)
    obs = np.random.default_rng(71).integers(0, M, size=(B, T)).astype(odt)
    o64 = obs.astype(np.int64)
    counts, W, sc = engine.estep_discrete_tc(h._params(), h._prep(obs), stash_out=True)
    W, sc, c = W.cpu().numpy(), sc.cpu().numpy(), counts.host()
    alpha, scale = O.forward(pi, A, O.emissions_discrete(Bm, o64))
    np.testing.assert_allclose(sc, scale, rtol=2e-3)
    np.testing.assert_allclose(W / sc[..., None], alpha, rtol=TC_TERM_RTOL, atol=1e-6)
    stats = None
    for b in range(B):
        s_ = O.sequence_stats_discrete(pi, A, Bm, o64[b])
        stats = s_ if stats is None else {k: stats[k] + s_[k] for k in s_}
    assert c["nseq"] == B and c["flags"] == 0
    assert np.isclose(c["loglik"], stats["loglik"], rtol=2e-5)
    np.testing.assert_allclose(c["pi"], stats["pi"], rtol=TC_TERM_RTOL, atol=1e-7)
    np.testing.assert_allclose(c["emis_den"], stats["emis_den"], rtol=2e-3)
    np.testing.assert_allclose(c["gam"], stats["gam"], rtol=2e-3, atol=2e-6 * stats["emis_den"].max())
    np.testing.assert_allclose(c["emis_num"], stats["emis_num"], rtol=TC_TERM_RTOL, atol=1e-7 * stats["emis_num"].max())
    if T > 1:
        np.testing.assert_allclose(c["xi_raw"] * A, stats["xi"], rtol=TC_TERM_RTOL, atol=1e-7 * stats["xi"].max())
    else:
        assert not c["xi_raw"].any()


def test_tensor_core_estep_many_tiles_vs_simt(K):
    """More tiles than CTAs (each CTA walks several tiles) and chunking: the tf32 statistics against
    the float32 SIMT path, which is itself pinned to the oracle above.  Sums over ~1e3..1e5 terms
    agree far tighter than single terms."""
    N, M, B, T = 128, 32, 148 * 128 * 2 + 300, 6
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=75)
    obs = np.random.default_rng(76).integers(0, M, size=(B, T)).astype(np.uint8)
    ref = h.estep_batch(obs, tensor_cores=False).host()
    for chunk in (None, 128 * 100):
        c = h.estep_batch(obs, tensor_cores=True, chunk=chunk).host()
        assert c["nseq"] == B and c["flags"] == 0
        assert np.isclose(c["loglik"], ref["loglik"], rtol=1e-5)
        for key in ("pi", "gam", "emis_den", "emis_num", "xi_raw"):
            # xi_raw carries a common factor of about 1 - 3.6e-4 (fp32 stash words truncated to tf32 by MMA2; it cancels in
            # the M-step's row normalisation): its bound is 1e-3, the normalised rows are checked below
            np.testing.assert_allclose(c[key], ref[key], rtol=1e-3 if key == "xi_raw" else 5e-4, err_msg=key)
        np.testing.assert_allclose(c["xi_raw"] / c["xi_raw"].sum(axis=1, keepdims=True),
                                   ref["xi_raw"] / ref["xi_raw"].sum(axis=1, keepdims=True), rtol=3e-4)


@pytest.mark.parametrize("B,T,M", [(300, 12, 16), (148 * 128 + 77, 9, 32)])
def test_tensor_core_estep_ragged_batches(K, B, T, M):
    """Ragged batches on the fused tensor-core E-step (rows padded to T with ARBITRARY symbols): the statistics are
    those of the unpadded sequences -- against the float32 SIMT path (itself pinned to the oracle on ragged batches)
    at the tf32 bounds of the equal-length test, and against the oracle's per-sequence statement on a sub-sample."""
    N = 128
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=81)
    rng = np.random.default_rng(82 + B)
    obs = rng.integers(0, M, size=(B, T)).astype(np.uint8)
    lengths = rng.integers(1, T + 1, size=B)
    lengths[:4] = [1, T, 2, T - 1]
    for b in range(B):
        obs[b, lengths[b]:] = rng.integers(0, 256)             # padding: anything, including out-of-alphabet symbols
    ref = h.estep_batch(obs, lengths=lengths, tensor_cores=False).host()
    c = h.estep_batch(obs, lengths=lengths, tensor_cores=True).host()
    assert c["nseq"] == B and c["flags"] == 0
    assert np.isclose(c["loglik"], ref["loglik"], rtol=1e-5)
    for key in ("pi", "gam", "emis_den", "emis_num", "xi_raw"):
        np.testing.assert_allclose(c[key], ref[key], rtol=1e-3 if key == "xi_raw" else 5e-4,
                                   atol=1e-6 * np.abs(ref[key]).max(), err_msg=key)
    # the same statistics from the oracle, sequence by sequence, on the first 40 sequences
    sub = slice(0, 40)
    c40 = h.estep_batch(obs[sub], lengths=lengths[sub], tensor_cores=True).host()
    tot = None
    for b in range(40):
        st = O.batch_stats_discrete(pi, A, Bm, obs[b:b + 1, :lengths[b]].astype(np.int64))
        tot = st if tot is None else {k: tot[k] + st[k] for k in st}
    assert np.isclose(c40["loglik"], tot["loglik"], rtol=2e-5)
    np.testing.assert_allclose(c40["emis_den"], tot["emis_den"], rtol=2e-3)
    np.testing.assert_allclose(c40["pi"], tot["pi"], rtol=2e-3, atol=1e-9)
    np.testing.assert_allclose(c40["xi_raw"] * A, tot["xi"], rtol=1e-2, atol=1e-6 * tot["xi"].max())


@pytest.mark.parametrize("N", [8, 16, 65, 100])
def test_tensor_core_estep_fewer_than_128_states(K, N):
    """8 <= N < 128 on the fused tensor-core E-step (padded with unreachable states): statistics against the float32
    SIMT path at the tf32 bounds, ragged batch included; one Baum-Welch pass leaves valid distributions."""
    M, B, T = 24, 148 * 128 + 50, 8
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=90 + N)
    rng = np.random.default_rng(91 + N)
    obs = rng.integers(0, M, size=(B, T)).astype(np.uint8)
    for lengths in (None, rng.integers(1, T + 1, size=B)):
        ref = h.estep_batch(obs, lengths=lengths, tensor_cores=False).host()
        c = h.estep_batch(obs, lengths=lengths, tensor_cores=True).host()
        assert c["nseq"] == B and c["flags"] == 0
        assert np.isclose(c["loglik"], ref["loglik"], rtol=1e-5)
        for key in ("pi", "gam", "emis_den", "emis_num", "xi_raw"):
            assert c[key].shape == ref[key].shape
            # xi_raw: the stash rows reach MMA2 as fp32 words, which the tensor core TRUNCATES to tf32 -- a common factor
            # of about 1 - 3e-4 on every entry (it cancels in the M-step's row normalisation)
            np.testing.assert_allclose(c[key], ref[key], rtol=1e-3 if key == "xi_raw" else 5e-4,
                                       atol=1e-6 * np.abs(ref[key]).max(), err_msg=key)
        xr = c["xi_raw"] / c["xi_raw"].sum(axis=1, keepdims=True)
        np.testing.assert_allclose(xr, ref["xi_raw"] / ref["xi_raw"].sum(axis=1, keepdims=True), rtol=3e-4)
    h.do_pass_batch(obs, tensor_cores=True)
    assert np.allclose(h.transitionMatrix.sum(axis=1), 1.0) and np.allclose(h.emission_matrix().sum(axis=1), 1.0)


def test_tensor_core_do_pass_batch(K):
    """One Baum-Welch pass through the public API with tensor_cores=True: re-estimated parameters
    against the oracle's M-step at the stated tf32 bound; they stay valid distributions."""
    N, M, B, T = 128, 48, 512, 64
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=77)
    obs = np.random.default_rng(78).integers(0, M, size=(B, T))
    stats = O.batch_stats_discrete(pi, A, Bm, obs)
    rpi, rA, rB = O.mstep_discrete(stats)
    ll = h.do_pass_batch(obs, tensor_cores=True)
    assert np.isclose(ll, stats["loglik"], rtol=2e-5)
    np.testing.assert_allclose(h.transitionMatrix, rA, rtol=2e-3)
    np.testing.assert_allclose(h.emission_matrix(), rB, rtol=2e-3)
    np.testing.assert_allclose(h.initialProbabilities, rpi, rtol=2e-3)
    h.sanity_check()


def test_train_batch_matches_oracle_iterations(K):
    """Several batched passes (float64) track the oracle's multi-sequence definition pass by pass.
    (The reference's weighting is not textbook EM, so the likelihood need not increase.)"""
    N, M, B, T = 6, 8, 40, 64
    teacher, _ = random_discrete(K, N, M, seed=41)
    np.random.seed(42)
    obs = np.array([teacher.simulate(T).emissions for _ in range(B)])
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=43)
    hist = h.train_batch(obs, iterations=4, dtype="float64")
    assert len(hist) == 4
    ref_hist = []
    for _ in range(4):
        stats = O.batch_stats_discrete(pi, A, Bm, obs)
        ref_hist.append(stats["loglik"])
        pi, A, Bm = O.mstep_discrete(stats)
    np.testing.assert_allclose(hist, ref_hist, rtol=1e-9)
    np.testing.assert_allclose(h.transitionMatrix, A, rtol=1e-7, atol=1e-14)
    np.testing.assert_allclose(h.emission_matrix(), Bm, rtol=1e-7, atol=1e-14)
    np.testing.assert_allclose(h.initialProbabilities, pi, rtol=1e-7, atol=1e-14)
    h.sanity_check()


def test_config1_full_length_golden(K):
    """BASELINE config 1 at full length (one sequence, T = 1000) through the reference-named single-sequence methods against
    the real reference's outputs (tests/golden/c1_golden.npz): float64 recursions to 1e-9, Viterbi path and log-probability
    bit-exact over 1000 steps, the bug-compatible forward_probability."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_golden.npz"))
    h = discrete_model(K, g["pi"], g["A"], g["B"])
    obs = g["obs"]
    alpha, scale = h.forward(list(obs))
    beta = h.backward(list(obs), scale)
    np.testing.assert_allclose(alpha, g["alpha"], rtol=1e-9)
    np.testing.assert_allclose(scale, g["scale"], rtol=1e-11)
    np.testing.assert_allclose(beta, g["beta"], rtol=1e-8)
    np.testing.assert_allclose(h.gamma(alpha, beta, scale), g["gamma"], rtol=1e-8)
    assert np.isclose(h.forward_probability(list(obs)), g["forward_probability"], rtol=1e-11)
    path, logp = h.viterbi_path(obs)
    assert path.dtype == obs.dtype and np.array_equal(path, g["path"]) and logp == g["logp"]


# ------------------------------------------------------------------------- errors / edges
def test_error_behaviour(K):
    h, _ = random_discrete(K, 4, 5, seed=0)
    with pytest.raises(IndexError):
        h.forward_batch(np.zeros((2, 3), dtype=np.float32))    # float symbols
    with pytest.raises(ValueError):
        h.forward_batch(np.zeros((2, 3), dtype=np.int64), lengths=[1, 4])
    with pytest.raises(ValueError):
        h.forward_batch(np.zeros((3,), dtype=np.int64))
    big = K.DiscreteHMM(1025, 2)
    from kerehmm_b200._lib import KhmmError
    with pytest.raises(KhmmError):
        big.forward_batch(np.zeros((1, 4), dtype=np.int64))
    # empty batch is fine
    a, s, ll = h.forward_batch(np.zeros((0, 7), dtype=np.int64))
    assert a.shape == (0, 7, 4) and ll.shape == (0,)
    p, lp = h.viterbi_batch(np.zeros((0, 7), dtype=np.int64))
    assert p.shape == (0, 7)
    # T == 1 do_pass: 0/0 in the transition update -> the reference's sanity assertion (Q13)
    with pytest.raises((AssertionError, ZeroDivisionError, ValueError)):
        h.do_pass([1])


def test_parameters_are_reread_on_every_call(K):
    """Callers mutate the attributes between calls (discrete_hmm_test.py:52-55)."""
    h = K.DiscreteHMM(3, 3)
    a1, _ = h.forward([0, 1, 2])
    h.transitionMatrix[:] = np.array([[.8, .1, .1], [.1, .8, .1], [.1, .1, .8]])
    h.emissionDistributions[0].probabilities[:] = [.9, .05, .05]
    a2, c2 = h.forward([0, 1, 2])
    assert not np.allclose(a1, a2)
    Bm = h.emission_matrix()
    ra, rc = O.forward(h.initialProbabilities, h.transitionMatrix, O.emissions_discrete(Bm, np.array([0, 1, 2])))
    np.testing.assert_allclose(a2, ra, rtol=1e-12)


# ---------------------------------------------- full-size properties (BASELINE.json shapes)
def test_viterbi_config3_shape_properties(K):
    """N=64, M=256, T=4096 at a reduced batch: the returned log-probability must equal, bit for
    bit, the score of the returned path recomputed in the DP's own order; and the oracle agrees on
    a sub-sample."""
    N, M, B, T = 64, 256, 512, 4096
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=3)
    obs = np.random.default_rng(1003).integers(0, M, size=(B, T)).astype(np.uint8)
    path, logp = h.viterbi_batch(obs, path_dtype="uint8")
    assert path.dtype == np.uint8
    with np.errstate(divide="ignore"):
        lpi, lA, lB = np.log(pi), np.log(A), np.log(Bm)
    p = path.astype(np.int64)
    o = obs.astype(np.int64)
    score = lpi[p[:, 0]] + lB[p[:, 0], o[:, 0]]
    for t in range(1, T):
        score = (score + lA[p[:, t - 1], p[:, t]]) + lB[p[:, t], o[:, t]]
    assert np.array_equal(score, logp)
    rp, rl = O.viterbi(pi, A, Bm, o[:8])
    assert np.array_equal(p[:8], rp) and np.array_equal(logp[:8], rl)


def test_viterbi_config3_full_size_two_kernels_agree(K):
    """BASELINE config 3 at FULL size (65,536 sequences x 4,096 steps, N = 64, M = 256): the float32-screened kernel
    and the plain float64 register-tile kernel -- two independent implementations of abstract_hmm.py:203-239 -- must
    return the same paths and the same float64 log-probabilities, bit for bit; the oracle agrees on a sub-sample."""
    import os
    import torch
    N, M, B, T = 64, 256, 65536, 4096
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=3)
    obs = torch.from_numpy(np.random.default_rng(1003).integers(0, M, size=(B, T)).astype(np.uint8)).cuda()
    from kerehmm_b200 import _lib
    try:
        _lib.call("khmm_debug_switch", b"KHMM_VIT_SCREEN", 1)
        p1, l1 = h.viterbi_batch(obs, path_dtype=torch.uint8)
        _lib.call("khmm_debug_switch", b"KHMM_VIT_SCREEN", 0)
        p0, l0 = h.viterbi_batch(obs, path_dtype=torch.uint8)
    finally:
        _lib.call("khmm_debug_switch", b"KHMM_VIT_SCREEN", -1)
    assert torch.equal(p1, p0) and torch.equal(l1, l0)
    with np.errstate(divide="ignore"):
        rp, rl = O.viterbi(pi, A, Bm, obs[:4].cpu().numpy().astype(np.int64))
    assert np.array_equal(p1[:4].cpu().numpy(), rp) and np.array_equal(l1[:4].cpu().numpy(), rl)


@pytest.mark.parametrize("mode,tol", [(True, 5e-6), ("bf16", 4e-5)])
def test_fwdbwd_config5_full_size_properties(K, mode, tol):
    """BASELINE config 5 at FULL size (GaussianHMM N = 1024, B = 4096, T = 1024) on the whole-recursion tensor-core
    kernel: the forward and the (independent, self-scaled) backward recursion must give the same log-likelihood for
    every sequence, and the float64 oracle agrees on a sub-sample (stated bounds: tf32 5e-6, bf16 4e-5 relative)."""
    import torch
    N, B, T = 1024, 4096, 1024
    np.random.seed(5)
    pi, A, mean, sd = O.random_gaussian_model(N)
    h = gaussian_model(K, pi, A, mean, sd)
    rng = np.random.default_rng(1005)
    x = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
    ll, llb = h.log_likelihood_batch(torch.from_numpy(x).cuda(), tensor_cores=mode)
    ll, llb = ll.cpu().numpy(), llb.cpu().numpy()
    assert np.isfinite(ll).all() and np.isfinite(llb).all()
    np.testing.assert_allclose(ll, llb, rtol=tol)
    ref = O.loglik(O.forward(pi, A, O.emissions_gaussian_1d(mean, sd, x[:8].astype(np.float64)))[1])
    np.testing.assert_allclose(ll[:8], ref, rtol=tol)
    np.testing.assert_allclose(llb[:8], ref, rtol=tol)


def test_baumwelch_config4_full_size_properties(K):
    """BASELINE config 4 at its FULL per-GPU size (DiscreteHMM N = 128, M = 1024, 131,072 sequences x 512 steps):
    three independent implementations of the scaled forward recursion -- the fused tensor-core E-step, the
    whole-recursion tensor-core log-likelihood kernel and the float32 SIMT forward kernel -- must agree on the
    log-likelihood; the E-step's counts satisfy the identities of its definition."""
    N, M, B, T = 128, 1024, 131072, 512
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=4)
    obs = np.random.default_rng(1004).integers(0, M, size=(B, T)).astype(np.uint16)
    c = h.estep_batch(obs, tensor_cores=True).host()
    assert c["nseq"] == B and c["flags"] == 0
    np.testing.assert_allclose(c["pi"].sum(), B, rtol=1e-5)                        # every smooth1d(gamma_0) sums to 1
    np.testing.assert_allclose(c["emis_num"].sum(axis=1), c["emis_den"], rtol=1e-6)
    assert (c["xi_raw"] >= 0).all() and np.isfinite(c["xi_raw"]).all()
    ll_simt, _ = h.log_likelihood_batch(obs, tensor_cores=False)
    ll_tc, llb_tc = h.log_likelihood_batch(obs, tensor_cores=True)
    np.testing.assert_allclose(ll_tc, ll_simt, rtol=5e-5)
    np.testing.assert_allclose(llb_tc, ll_simt, rtol=5e-5)
    np.testing.assert_allclose(c["loglik"], ll_simt.sum(), rtol=2e-5)
    ref = O.loglik(O.forward(pi, A, O.emissions_discrete(Bm, obs[:4].astype(np.int64)))[1])
    np.testing.assert_allclose(ll_simt[:4], ref, rtol=F32_RTOL)


@pytest.mark.parametrize("N,B,T", [(128, 1, 1), (256, 3, 2), (128, 1, 7), (512, 129, 3)])
def test_tensor_core_loglik_degenerate_shapes(K, N, B, T):
    """The whole-recursion kernel at its edges: a single time step (no recursion job at all), two steps, a single
    sequence (one row tile with 127 padded rows), one sequence more than a row tile."""
    import torch
    np.random.seed(400 + N + T)
    pi, A, mean, sd = O.random_gaussian_model(N)
    h = gaussian_model(K, pi, A, mean, sd)
    rng = np.random.default_rng(401 + B)
    x = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
    ref = O.loglik(O.forward(pi, A, O.emissions_gaussian_1d(mean, sd, x.astype(np.float64)))[1])
    # operands rounded to tf32 / bf16: ~2^-11 / 2^-8 per emission value, and with sd = 1 only a few states carry a step's
    # probability, so the ABSOLUTE error of log P is ~1e-3 / ~1e-2 per step at worst -- visible on these tiny T
    for mode, atol in ((True, 3e-3 * T), ("bf16", 2e-2 * T)):
        ll, llb = h.log_likelihood_batch(torch.from_numpy(x).cuda(), tensor_cores=mode)
        assert np.isfinite(ll.cpu().numpy()).all() and np.isfinite(llb.cpu().numpy()).all()
        np.testing.assert_allclose(ll.cpu().numpy(), ref, rtol=0, atol=atol)
        np.testing.assert_allclose(llb.cpu().numpy(), ref, rtol=0, atol=atol)


@pytest.mark.parametrize("N,B,T", [(256, 200, 20), (1024, 300, 12), (128, 70, 33)])
@pytest.mark.parametrize("mode,tol", [(True, 5e-5), ("bf16", 3e-4)])
def test_tensor_core_loglik_ragged_batches(K, N, B, T, mode, tol):
    """Ragged batches on the whole-recursion tensor-core kernel (rows padded to T with garbage, NaN included): the
    forward and the backward log-likelihood of every sequence against the float64 oracle run on the unpadded sequence."""
    import torch
    np.random.seed(500 + N)
    pi, A, mean, sd = O.random_gaussian_model(N)
    h = gaussian_model(K, pi, A, mean, sd)
    rng = np.random.default_rng(501 + B)
    x = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
    lengths = rng.integers(1, T + 1, size=B)
    lengths[:4] = [1, T, 2, T - 1]
    for b in range(B):
        x[b, lengths[b]:] = np.nan if b % 3 == 0 else 1e6 * rng.standard_normal()
    ll, llb = h.log_likelihood_batch(torch.from_numpy(x).cuda(), lengths=lengths, tensor_cores=mode)
    ll, llb = ll.cpu().numpy(), llb.cpu().numpy()
    ref = np.array([O.loglik(O.forward(pi, A, O.emissions_gaussian_1d(mean, sd, x[b, :lengths[b]].astype(np.float64)))[1])
                    for b in range(B)])
    # short sequences: the absolute error of one step's rounded emission values dominates (see the degenerate-shape test)
    atol = (3e-3 if mode is True else 2e-2)
    np.testing.assert_allclose(ll, ref, rtol=tol, atol=atol)
    np.testing.assert_allclose(llb, ref, rtol=tol, atol=atol)


@pytest.mark.parametrize("N,kind", [(100, "gauss"), (200, "gauss"), (1000, "gauss"), (65, "discrete"), (300, "discrete")])
def test_tensor_core_loglik_any_state_count(K, N, kind):
    """State spaces that are not a multiple of 128 run on the tensor-core kernel padded with unreachable states
    (start probability 0, no transitions, emission 0): likelihoods unchanged, ragged batches included."""
    import torch
    B, T = 150, 24
    rng = np.random.default_rng(600 + N)
    lengths = rng.integers(1, T + 1, size=B)
    if kind == "gauss":
        np.random.seed(601 + N)
        pi, A, mean, sd = O.random_gaussian_model(N)
        h = gaussian_model(K, pi, A, mean, sd)
        x = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
        E = O.emissions_gaussian_1d(mean, sd, x.astype(np.float64))
    else:
        h, (pi, A, Bm) = random_discrete(K, N, 40, seed=602 + N)
        x = rng.integers(0, 40, size=(B, T)).astype(np.uint8)
        E = O.emissions_discrete(Bm, x)
    ref = O.loglik(O.forward(pi, A, E)[1])
    ll, llb = h.log_likelihood_batch(torch.from_numpy(x).cuda(), tensor_cores=True)
    assert llb is not None                                           # it really took the tensor-core path
    np.testing.assert_allclose(ll.cpu().numpy(), ref, rtol=5e-5)
    np.testing.assert_allclose(llb.cpu().numpy(), ref, rtol=5e-5)
    ref_r = np.array([O.loglik(O.forward(pi, A, E[b, :lengths[b]])[1]) for b in range(B)])
    ll, llb = h.log_likelihood_batch(torch.from_numpy(x).cuda(), lengths=lengths, tensor_cores=True)
    np.testing.assert_allclose(ll.cpu().numpy(), ref_r, rtol=5e-5, atol=3e-3)
    np.testing.assert_allclose(llb.cpu().numpy(), ref_r, rtol=5e-5, atol=3e-3)


def test_fwdbwd_config2_shape_properties(K):
    """N=8 Gaussian, T=2048 at a reduced batch: forward and backward likelihoods agree, and the
    oracle agrees on a sub-sample."""
    N, B, T = 8, 2048, 2048
    np.random.seed(2)
    pi, A, mean, sd = O.random_gaussian_model(N)
    h = gaussian_model(K, pi, A, mean, sd)
    rng = np.random.default_rng(1002)
    obs = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
    ll, llb = h.log_likelihood_batch(obs)
    np.testing.assert_allclose(ll, llb, rtol=2e-6)
    E = O.emissions_gaussian_1d(mean, sd, obs[:16].astype(np.float64))
    ref = O.loglik(O.forward(pi, A, E)[1])
    np.testing.assert_allclose(ll[:16], ref, rtol=F32_RTOL)
    gam, ll3 = h.posterior_batch(obs[:64])
    np.testing.assert_allclose(gam.sum(axis=2), 1.0, rtol=1e-4)
    np.testing.assert_allclose(ll3, ll[:64], rtol=F32_RTOL)


def test_loglik_host_batch_pipelined_path_matches_device_path(K):
    """A large host batch takes the copy/compute-overlapped path (row blocks on a side stream); its results
    equal the device-resident path bit for bit (same kernel, same rows)."""
    import torch
    np.random.seed(91)
    pi, A, mean, sd = O.random_gaussian_model(8)
    h = gaussian_model(K, pi, A, mean, sd)
    rng = np.random.default_rng(92)
    x = (mean[rng.integers(0, 8, size=(5000, 1024))] + rng.standard_normal((5000, 1024))).astype(np.float32)
    ll_h, llb_h = h.log_likelihood_batch(x)                                   # 20 MB host array -> pipelined
    ll_d, llb_d = h.log_likelihood_batch(torch.from_numpy(x).cuda())          # device tensor -> plain
    assert isinstance(ll_h, np.ndarray) and ll_h.shape == (5000,)
    assert np.array_equal(ll_h, ll_d.cpu().numpy()) and np.array_equal(llb_h, llb_d.cpu().numpy())
    ll_p, _ = h.log_likelihood_batch(torch.from_numpy(x).pin_memory(), backward_check=False)
    assert np.array_equal(ll_p, ll_h)


def test_tensor_core_loglik_persistent_tiles(K):
    """N = 512 with more (row tile, column tile, direction) jobs than SMs (several jobs per CTA and step, TMEM
    accumulators double-buffered) against the float64 oracle."""
    import torch
    N, B, T = 512, 38 * 128, 6
    np.random.seed(95)
    pi, A, mean, sd = O.random_gaussian_model(N)
    h = gaussian_model(K, pi, A, mean, sd)
    rng = np.random.default_rng(96)
    x = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
    ll, llb = h.log_likelihood_batch(torch.from_numpy(x).cuda())
    ref = O.fwdbwd_loglik_batch(pi, A, O.emissions_gaussian_1d(mean, sd, x.astype(np.float64)))
    ref = ref[0] if isinstance(ref, tuple) else ref
    np.testing.assert_allclose(ll.cpu().numpy(), ref, rtol=5e-5)
    np.testing.assert_allclose(llb.cpu().numpy(), ref, rtol=5e-5)


def test_tensor_core_loglik_bf16_operands(K):
    """bf16-operand mode of the whole-recursion kernel (kind::f16, fp32 accumulation): stated bound 1e-4 relative
    on a short sequence (the absolute error grows like sqrt(T), log P like T: 7e-5 measured at T = 24, 2e-5 at
    T = 64, 7e-6 at T = 1024), and the fallback to tf32 for shapes the mode does not cover."""
    import torch
    N, B, T = 512, 38 * 128, 24
    np.random.seed(97)
    pi, A, mean, sd = O.random_gaussian_model(N)
    h = gaussian_model(K, pi, A, mean, sd)
    rng = np.random.default_rng(98)
    x = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
    ll, llb = h.log_likelihood_batch(torch.from_numpy(x).cuda(), tensor_cores="bf16")
    ref = O.loglik(O.forward(pi, A, O.emissions_gaussian_1d(mean, sd, x[:256].astype(np.float64)))[1])
    np.testing.assert_allclose(ll.cpu().numpy()[:256], ref, rtol=1e-4)
    np.testing.assert_allclose(llb.cpu().numpy()[:256], ref, rtol=1e-4)
    ll_tf, _ = h.log_likelihood_batch(torch.from_numpy(x).cuda())
    assert not np.array_equal(ll.cpu().numpy(), ll_tf.cpu().numpy())          # it really is another arithmetic
    small, _ = h.log_likelihood_batch(torch.from_numpy(x[:256]).cuda(), tensor_cores="bf16")   # few jobs: 128-wide tiles
    np.testing.assert_allclose(small.cpu().numpy(), ref, rtol=1e-4)
    np.testing.assert_array_equal(small.cpu().numpy(), ll.cpu().numpy()[:256])  # a row does not depend on the batch shape


@pytest.mark.gpu
@pytest.mark.parametrize("N,B,T", [(128, 200, 12), (384, 200, 12), (1024, 512, 8), (256, 70, 9)])
@pytest.mark.parametrize("mode", [True, "bf16"])
def test_tensor_core_loglik_small_batches(K, N, B, T, mode):
    """Whole-recursion kernel on shapes with fewer jobs than SMs (a strong-scaling shard of config 5 is B = 512),
    N not a multiple of 256 (128-wide column tiles) and ragged last row tiles, against the float64 oracle."""
    import torch
    np.random.seed(101 + N)
    pi, A, mean, sd = O.random_gaussian_model(N)
    h = gaussian_model(K, pi, A, mean, sd)
    rng = np.random.default_rng(102 + B)
    x = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
    ll, llb = h.log_likelihood_batch(torch.from_numpy(x).cuda(), tensor_cores=mode)
    ref = O.loglik(O.forward(pi, A, O.emissions_gaussian_1d(mean, sd, x.astype(np.float64)))[1])
    tol = 2e-4 if mode == "bf16" else 5e-5
    np.testing.assert_allclose(ll.cpu().numpy(), ref, rtol=tol)
    np.testing.assert_allclose(llb.cpu().numpy(), ref, rtol=tol)


# ----------------------------------------------- full-covariance Gaussian emissions fused into the recursions (round 2)
@pytest.mark.parametrize("D,dtype,tol", [(2, "float64", 1e-10), (3, "float32", 3e-5),
                                          (8, "float64", 1e-10), (9, "float64", 1e-10)])
def test_full_covariance_emissions_fused_in_the_recursions(K, D, dtype, tol):
    """GaussianDistribution with a covariance matrix (distribution.py:153) on a ragged batch: forward / backward with the
    density evaluated inside the recursion kernels (D <= 8; D = 9 takes the dense emission matrix) against the oracle's
    per-sequence forward / backward on SciPy-checked densities (abstract_hmm.py:250-285, 327-351)."""
    N, B, T = 5, 23, 40
    rng = np.random.default_rng(100 + D)
    np.random.seed(D)
    h = K.GaussianHMM(N, D, random_transitions=True)
    means = rng.normal(scale=3.0, size=(N, D))
    covs = np.empty((N, D, D))
    for j in range(N):
        Q = rng.normal(size=(D, D))
        covs[j] = Q @ Q.T / D + 0.5 * np.eye(D)
    h.emissionDistributions = [K.GaussianDistribution(D, mean=means[j], variance=covs[j]) for j in range(N)]
    x = means[rng.integers(0, N, size=(B, T))] + rng.normal(size=(B, T, D))
    lengths = rng.integers(1, T + 1, size=B)
    lengths[0] = T
    alpha, scale, ll = h.forward_batch(x, lengths=lengths, dtype=dtype)
    beta = h.backward_batch(x, scale, lengths=lengths, dtype=dtype)
    gam, ll2 = h.posterior_batch(x, lengths=lengths, dtype=dtype)
    pi, A = h.initialProbabilities, h.transitionMatrix
    for b in range(B):
        L = int(lengths[b])
        E = O.emissions_gaussian_full(means, covs, x[b, :L])
        ra, rc = O.forward(pi, A, E)
        rb = O.backward(A, E, rc)
        np.testing.assert_allclose(alpha[b, :L], ra, rtol=tol, atol=1e-30)
        np.testing.assert_allclose(scale[b, :L], rc, rtol=tol)
        np.testing.assert_allclose(beta[b, :L], rb, rtol=tol * 10, atol=1e-30)
        np.testing.assert_allclose(gam[b, :L], ra * rb, rtol=tol * 10, atol=1e-12)
        assert np.isclose(ll[b], np.log(rc).sum(), rtol=max(tol, 1e-12)) and np.isclose(ll2[b], ll[b], rtol=1e-12)
