"""GPU tests of the batched sampler (khmm_simulate_*).  The sampling LOGIC is pinned
to the real reference through golden vectors (states/emissions of `simulate` plus the MT19937 doubles it
consumed); the device generator itself (Philox4x32-10) is pinned by the published known-answer vectors;
its agreement with the reference's RNG stream is distributional only."""
import ctypes
import os

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


def _discrete(K, pi, A, Bm):
    h = K.DiscreteHMM(A.shape[0], Bm.shape[1])
    h.initialProbabilities, h.transitionMatrix = pi.copy(), A.copy()
    for i, d in enumerate(h.emissionDistributions):
        d.probabilities = Bm[i].copy()
    return h


def test_philox4x32_10_known_answers(K):
    """Random123 known-answer vectors for philox4x32-10 (counter[4], key[2] -> output[4])."""
    import torch
    from kerehmm_b200 import _lib, engine
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    inp = np.array([c + k for c, k, _ in kat], dtype=np.uint32)
    d_in = torch.from_numpy(inp.view(np.int32)).cuda()
    d_out = torch.zeros((len(kat), 4), dtype=torch.int32, device="cuda")
    _lib.call("khmm_philox4x32_10", len(kat), d_in.data_ptr(), d_out.data_ptr(), engine.stream_ptr())
    got = d_out.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, np.array([o for _, _, o in kat], dtype=np.uint32))


@pytest.mark.parametrize("name", ["small", "wide"])
def test_simulate_matches_reference_golden(K, name):
    """With the doubles the reference's `simulate` consumed, the device returns the reference's states and
    emissions exactly."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "simulate_golden.npz"))
    h = _discrete(K, g[name + "_pi"], g[name + "_A"], g[name + "_B"])
    U = g[name + "_uniforms"]
    r = h.simulate_batch(U.shape[0], U.shape[1], uniforms=U)
    assert np.array_equal(r.states, g[name + "_states"]) and np.array_equal(r.emissions, g[name + "_emissions"])


def test_simulate_discrete_against_oracle_uniforms(K):
    np.random.seed(61)
    pi, A, Bm = O.random_discrete_model(128, 1024)
    U = np.random.default_rng(62).random((40, 30, 3))
    U[0, 0] = 0.0                               # edge: u = 0 picks the first state / symbol with p > 0
    U[1, 0] = np.nextafter(1.0, 0.0)            # edge: largest u picks the last
    r = _discrete(K, pi, A, Bm).simulate_batch(40, 30, uniforms=U)
    s, e = O.simulate_from_uniforms(pi, A, U, Bm=Bm)
    assert np.array_equal(r.states, s) and np.array_equal(r.emissions, e)


def test_simulate_gaussian_against_oracle_uniforms(K):
    np.random.seed(63)
    pi, A, mean, sd = O.random_gaussian_model(8)
    sd = sd * np.linspace(0.5, 2.0, 8)
    h = K.GaussianHMM(8, 1)
    h.initialProbabilities, h.transitionMatrix = pi.copy(), A.copy()
    h.emissionDistributions = [K.GaussianDistribution(1, mean=float(m), variance=float(s)) for m, s in zip(mean, sd)]
    U = np.random.default_rng(64).random((16, 50, 3))
    r = h.simulate_batch(16, 50, uniforms=U)
    s, v = O.simulate_from_uniforms(pi, A, U, mean=mean, sd=sd)
    assert np.array_equal(r.states, s)
    np.testing.assert_allclose(r.emissions, v, rtol=1e-12, atol=1e-12)


def test_simulate_philox_statistics_and_determinism(K):
    """Counter-based stream: same seed -> same sequences whatever the batch size; other seed -> other
    sequences; empirical start / transition / emission frequencies within 5 sigma of the model."""
    np.random.seed(65)
    pi, A, Bm = O.random_discrete_model(4, 6)
    h = _discrete(K, pi, A, Bm)
    B, T = 4000, 300
    r = h.simulate_batch(B, T, seed=1234)
    r2 = h.simulate_batch(100, T, seed=1234)
    assert np.array_equal(r.states[:100], r2.states) and np.array_equal(r.emissions[:100], r2.emissions)
    assert not np.array_equal(h.simulate_batch(100, T, seed=1235).states, r2.states)
    s, e = r.states, r.emissions
    start = np.bincount(s[:, 0], minlength=4) / B
    assert np.all(np.abs(start - pi) < 5 * np.sqrt(pi * (1 - pi) / B))
    for i in range(4):
        frm = s[:, :-1] == i
        n = frm.sum()
        trans = np.array([(frm & (s[:, 1:] == j)).sum() for j in range(4)]) / n
        assert np.all(np.abs(trans - A[i]) < 5 * np.sqrt(A[i] * (1 - A[i]) / n + 1e-12))
        at = s == i
        emis = np.bincount(e[at], minlength=6) / at.sum()
        assert np.all(np.abs(emis - Bm[i]) < 5 * np.sqrt(Bm[i] * (1 - Bm[i]) / at.sum() + 1e-12))


def test_simulate_gaussian_philox_moments(K):
    h = K.GaussianHMM(3, 1)
    h.emissionDistributions = [K.GaussianDistribution(1, mean=m, variance=s) for m, s in ((-5.0, 0.5), (0.0, 1.0), (7.0, 2.0))]
    r = h.simulate_batch(2000, 200, seed=9)
    for i, (m, sd) in enumerate(((-5.0, 0.5), (0.0, 1.0), (7.0, 2.0))):
        x = r.emissions[r.states == i]
        assert abs(x.mean() - m) < 6 * sd / np.sqrt(x.size) and abs(x.std() - sd) < 0.02 * sd
    assert r.emissions.dtype == np.float64 and r.states.dtype == np.int32
