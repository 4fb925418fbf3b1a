"""Host-side mirror of the reference API: constructors, RNG draw order, sanity checks, smoothing,
fixtures, distributions -- everything that runs without a GPU -- plus the loud failure when the
CUDA path is unavailable (there is no CPU fallback)."""
import numpy as np
import pytest

import kerehmm_b200 as K
from conftest import gcase
from kerehmm_b200 import util


def test_constructors_draw_like_the_reference(golden):
    c = gcase(golden, "CTOR_D")
    np.random.seed(32)
    h = K.DiscreteHMM(4, 6, random_transitions=True, random_emissions=True)
    assert np.array_equal(h.initialProbabilities, c["pi"])
    assert np.array_equal(h.transitionMatrix, c["A"])
    assert np.array_equal(h.emission_matrix(), c["B"])
    assert h.alphabetSize == 6 and h.nStates == 4 and h.stateLabels[2] == "State_2"
    assert isinstance(h.emissionDistributions, np.ndarray)            # object ndarray (Q12)
    c = gcase(golden, "CTOR_G")
    np.random.seed(33)
    g = K.GaussianHMM(5, 1, random_transitions=True, random_emissions=True)
    assert isinstance(g.emissionDistributions, list)                  # plain list (Q12)
    assert np.array_equal([d.mean for d in g.emissionDistributions], c["mean"])
    assert np.array_equal([d.variance for d in g.emissionDistributions], c["sd"])
    assert np.array_equal(g.transitionMatrix, c["A"]) and np.array_equal(g.initialProbabilities, c["pi"])


def test_random_init_and_array_sizes():
    """kerehmm/test/discrete_hmm_test.py:37-46."""
    for _ in range(200):
        K.DiscreteHMM(3, 3, random_transitions=True, random_emissions=True)
    h = K.DiscreteHMM(3, 5)
    assert len(h.emissionDistributions) == 3
    assert h.transitionMatrix.shape == (3, 3) and h.initialProbabilities.shape == (3,)
    assert all(d.n == 5 for d in h.emissionDistributions)
    assert np.allclose(h.transitionMatrix, 1 / 3.) and np.allclose(h.emission_matrix(), 1 / 5.)
    with pytest.raises(AssertionError):
        K.DiscreteHMM(2, 2, state_labels=["a", "a"])
    assert h.l2s("State_1") == 1 and h.s2l(2) == "State_2"
    assert h.transition_probability(0, 1) == h.transitionMatrix[0, 1]
    assert h.initial_probability(1) == h.initialProbabilities[1]


def test_sanity_check_semantics():
    h = K.DiscreteHMM(3, 3)
    h.sanity_check()
    with pytest.raises(TypeError):
        h.sanity_check(sanitize=True)              # the Discrete override takes no arguments (Q13)
    h.transitionMatrix[1] = [.5, .5, .5]
    with pytest.raises(AssertionError):
        h.sanity_check()
    g = K.GaussianHMM(3, 1)
    g.transitionMatrix[1] = [.5, .5, .5]
    with pytest.raises(AssertionError):
        g.sanity_check()
    g.sanity_check(sanitize=True)
    assert np.allclose(g.transitionMatrix.sum(axis=1), 1)
    h = K.DiscreteHMM(3, 3)
    h.emissionDistributions[0].probabilities = np.array([.5, .5, .5])
    with pytest.raises(AssertionError):
        h.sanity_check()


def test_smooth_probabilities_matches_reference(golden):
    g = gcase(golden, "SMOOTH")
    x = g["in1"].copy()
    assert util.smooth_probabilities(x) is x and np.array_equal(x, g["out1"])
    assert np.array_equal(util.smooth_probabilities(g["in2"].copy()), g["out2"])
    with pytest.raises(ValueError):
        util.smooth_probabilities(g["in3"].copy())
    with pytest.raises(ZeroDivisionError):
        util.smooth_probabilities(np.zeros(4))
    with pytest.raises(ValueError):
        util.smooth_probabilities(np.zeros((2, 2, 2)))
    s = gcase(golden, "SIMPLEX")
    np.random.seed(31)
    assert np.array_equal(util.random_simplex(7), s["one_d"])
    assert np.array_equal(util.random_simplex(5, two_d=True), s["two_d"])
    assert util.DELTA_P == 1e-10 and util.CONVERGENCE_DELTA_LOG_LIKELIHOOD == 1e-5
    assert np.isclose(util.add_logs([np.log(.25), np.log(.5)]), np.log(.75))


def test_left_to_right_fixtures(golden):
    h = K.DiscreteHMM(3, 3)
    h.setup_strict_left_to_right(set_emissions=True)
    g = gcase(golden, "KAT_l2r_emis")
    assert np.array_equal(h.initialProbabilities, g["pi"]) and np.array_equal(h.transitionMatrix, g["A"])
    assert np.array_equal(h.emission_matrix(), g["B"])
    h = K.DiscreteHMM(4, 4)
    h.setup_left_to_right(set_emissions=True)
    g = gcase(golden, "KAT_l2r")
    assert np.array_equal(h.initialProbabilities, g["pi"]) and np.array_equal(h.transitionMatrix, g["A"])
    assert np.array_equal(h.emission_matrix(), g["B"])


def test_distributions(golden):
    d = K.DiscreteDistribution(3)
    assert np.allclose(d.probabilities, 1 / 3.) and d[1] == d.probabilities[1]
    assert d.b_coefficient(K.DiscreteDistribution(3)) == np.sqrt(np.sum([1. / 3 ** 2] * 3))   # distribution.py:47-59
    np.random.seed(1)
    assert 0 <= d.emit() < 3
    m = K.GaussianDistribution(2)
    assert np.array_equal(m.mean, [0., 0.]) and np.array_equal(m.variance, np.eye(2))
    assert np.isclose(m[(0, 0)], 0.15915494309189535) and np.isclose(m[(1, 2)], 0.013064233284684921)
    u = K.GaussianDistribution(1)
    assert u.mean == 0 and u.variance == 1
    assert np.isclose(u[0], 0.3989422804014327)
    g3 = gcase(golden, "G3")
    assert np.isclose(K.GaussianDistribution(1, mean=3., variance=4.)[5.0], g3["pdf_3_4_at_5"], rtol=1e-15)
    r = K.GaussianDistribution(2, random=True, lower_bounds=[1, 1], upper_bounds=[2, 2])
    assert all(1 <= v <= 2 for v in r.mean)
    with pytest.raises(ValueError):
        K.GaussianDistribution(2, mean=np.zeros(3))
    with pytest.raises(NotImplementedError):
        K.Distribution()[0]
    gm = gcase(golden, "GM")
    from scipy.stats import multivariate_normal
    dd = K.GaussianDistribution(2, mean=gm["means"][0], variance=gm["covs"][0])
    assert np.isclose(dd[gm["obs"][0]], multivariate_normal(gm["means"][0], gm["covs"][0]).pdf(gm["obs"][0]), rtol=1e-12)


def test_emission_probability_and_simulate():
    h = K.DiscreteHMM(3, 4, random_emissions=True)
    v = h.emission_probability(None, 2)
    assert v.shape == (3,) and v[1] == h.emissionDistributions[1].probabilities[2]
    assert h.emission_probability(1, 2) == v[1]
    np.random.seed(0)
    res = h.simulate(25)
    assert len(res.states) == 25 and len(res.emissions) == 25 and max(res.emissions) < 4
    assert res._fields == ("states", "emissions")
    g = K.GaussianHMM(2, 1, random_emissions=True)
    assert len(g.simulate(5).emissions) == 5


def test_state_alignment_helpers():
    a, b = K.DiscreteHMM(3, 3), K.DiscreteHMM(3, 3)
    a.initialProbabilities = np.array([.1, .7, .2])
    b.initialProbabilities = np.array([.2, .7, .1])
    assert a.compute_mapping_to(b) == (2, 1, 0)
    a.rearrange_like(b)
    assert np.array_equal(a.initialProbabilities, b.initialProbabilities)
    with pytest.raises(NotImplementedError):
        K.AbstractHMM(2).compute_emission_mapping_score(None, None)


def test_abstract_stubs():
    h = K.AbstractHMM(2)
    for call in (lambda: h.do_pass([0]), lambda: h.backward_probability([0]), lambda: h.initialize_parameters([0])):
        with pytest.raises(NotImplementedError):
            call()


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry fails loudly instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = K.DiscreteHMM(3, 3)
    for call in (lambda: h.forward([0, 1]), lambda: h.viterbi_path([0, 1]), lambda: h.do_pass([0, 1, 2]),
                 lambda: h.forward_batch(np.zeros((2, 3), dtype=np.int64)),
                 lambda: h.viterbi_batch(np.zeros((2, 3), dtype=np.int64)),
                 lambda: h.log_likelihood_batch(np.zeros((2, 3), dtype=np.int64))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) must print ONE JSON line with
    the contract's keys; the smallest workload keeps this a sub-second CPU test."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "transitions/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["value"] > 0 and d["steps"] == 2 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_bench_reference_arm_under_torchrun_prints_once():
    """The driver launches the reference arm like the GPU arm (torchrun for N > 1): rank 0 alone runs and prints, the
    other ranks exit 0 without work."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1", "--gpus", "2",
           "--steps", "1", "--warmup", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0


def test_state_alignment_exact_and_assignment_based():
    """compute_mapping_to: exhaustive search up to 8 states (the reference's own objective, abstract_hmm.py:63-94) and
    the Hungarian + swap-descent path above it; both recover a relabelled copy of a model, and on small models the
    assignment path reaches the exhaustive optimum."""
    import kerehmm_b200 as K
    rng = np.random.default_rng(11)
    for n, m in ((5, 7), (12, 9), (40, 16)):
        np.random.seed(100 + n)
        a = K.DiscreteHMM(n, m, random_transitions=True, random_emissions=True)
        perm = list(rng.permutation(n))
        b = K.DiscreteHMM(n, m)
        b.initialProbabilities = a.initialProbabilities[perm]
        b.transitionMatrix = a.transitionMatrix[perm, :][:, perm]
        b.emissionDistributions = a.emissionDistributions[perm]
        mapping = a.compute_mapping_to(b)
        assert list(mapping) == perm                      # b's state k is a's state perm[k]
        a.rearrange_like(b)
        assert np.allclose(a.transitionMatrix, b.transitionMatrix) and np.allclose(a.emission_matrix(), b.emission_matrix())
    # two unrelated 6-state models: the transition term makes this a quadratic assignment problem, so the assignment path
    # ends in a local optimum of the same objective -- close to, never below, the exhaustive optimum
    np.random.seed(7)
    x = K.DiscreteHMM(6, 5, random_transitions=True, random_emissions=True)
    y = K.DiscreteHMM(6, 5, random_transitions=True, random_emissions=True)
    exact = x.compute_mapping_to(y)
    approx = x.compute_mapping_to(y, exact_max_states=0)
    assert x.mapping_score(y, exact) - 1e-12 <= x.mapping_score(y, approx) <= 1.1 * x.mapping_score(y, exact)
    # the emission score of a permutation is the sum of the pair costs
    c = x._emission_pair_costs(y)
    assert np.isclose(x.compute_emission_mapping_score(y, list(exact)), sum(c[k, mk] for k, mk in enumerate(exact)))
